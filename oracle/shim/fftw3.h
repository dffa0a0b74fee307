/* oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * The FFTW3 single-precision calls the reference's lib/fft.cc makes (FFTW is not in the image):
 * plans over caller-owned buffers, executed by oracle/shim/shim_impl.cc (float butterflies with
 * double-precision twiddle generation; unnormalised like FFTW).  Wisdom and threading calls are no-ops. */
#pragma once
#include <cstdio>

typedef float fftwf_complex[2];
typedef struct clb_shim_fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

extern "C" {
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
}
static inline int fftwf_import_wisdom_from_file(FILE *) { return 1; }
static inline void fftwf_export_wisdom_to_file(FILE *) {}
static inline int fftwf_init_threads(void) { return 1; }
static inline void fftwf_plan_with_nthreads(int) {}
