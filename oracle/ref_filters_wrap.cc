/*
 * oracle/ref_filters_wrap.cc -- TEST INFRASTRUCTURE ONLY.
 * extern "C" entry points over the REFERENCE's own CPU filter classes: fft_filter_ccf
 * (lib/fft_filter.cc), fir_filter_ccf (lib/fir_filter.cc) and the fft_complex wrapper they sit on
 * (lib/fft.cc), compiled by oracle/Makefile from where they lie under /root/reference into
 * oracle/_ref/libref_filters.so, against the stand-in headers in oracle/shim/ for the libraries the image
 * lacks (VOLK, FFTW3, Boost, GNU Radio).  So the blocking, overlap-add bookkeeping, tap reversal /
 * alignment tables and decimation counters are the reference's code, executed here; only the innermost
 * library kernels (FFT butterflies, dot products) are the shim's.  Used to pin oracle_c.c's restatement of
 * these classes and to generate tests/golden/ref_filters.npz.  No reference source is copied.
 */
#include <complex>
#include <cstring>
#include <vector>

#include "fft_filter.h"
#include "fir_filter.h"

using gr::clenabled::fft_complex;
using gr::clenabled::fft_filter_ccf;
using gr::clenabled::fir_filter_ccf;

#define REF_API extern "C" __attribute__((visibility("default")))

/* fft_filter_ccf(decimation, taps, nthreads) (lib/fft_filter.cc:18-26) */
REF_API void *ref_fftfilt_create(int decimation, const float *taps, int ntaps)
{
    try {
        return new fft_filter_ccf(decimation, std::vector<float>(taps, taps + ntaps), 1);
    } catch (...) {
        return nullptr;
    }
}
REF_API void ref_fftfilt_destroy(void *h) { delete static_cast<fft_filter_ccf *>(h); }
/* compute_sizes (lib/fft_filter.cc:72-97) */
REF_API void ref_fftfilt_sizes(void *h, int *fftsize, int *nsamples)
{
    auto *f = static_cast<fft_filter_ccf *>(h);
    *fftsize = f->d_fftsize;
    *nsamples = f->d_nsamples;
}
/* set_taps (lib/fft_filter.cc:38-68): returns nsamples, resets the tail */
REF_API int ref_fftfilt_set_taps(void *h, const float *taps, int ntaps)
{
    return static_cast<fft_filter_ccf *>(h)->set_taps(std::vector<float>(taps, taps + ntaps));
}
/* filter(nitems, input, output) (lib/fft_filter.cc:129-175): nitems OUTPUT items, nitems*decimation inputs */
REF_API int ref_fftfilt_filter(void *h, int nitems, const float *in, float *out)
{
    return static_cast<fft_filter_ccf *>(h)->filter(nitems, reinterpret_cast<const gr_complex *>(in),
                                                    reinterpret_cast<gr_complex *>(out));
}
/* the transformed taps the filter multiplies by (d_xformed_taps, fftsize complex values) */
REF_API void ref_fftfilt_xformed_taps(void *h, float *out)
{
    auto *f = static_cast<fft_filter_ccf *>(h);
    std::memcpy(out, f->d_xformed_taps, sizeof(gr_complex) * f->d_fftsize);
}

/* fir_filter_ccf::filterN / filterNdec (lib/fir_filter.cc:235-258): in holds n*decimate + ntaps - 1 samples */
REF_API int ref_fir_ccf(const float *taps, int ntaps, const float *in, long n, int decimate, float *out)
{
    try {
        fir_filter_ccf f(decimate, std::vector<float>(taps, taps + ntaps));
        // the class rounds input pointers down to the VOLK alignment: hand it an aligned copy at every offset
        const size_t len = (size_t)n * decimate + ntaps;
        std::vector<gr_complex> buf(len + 8);
        gr_complex *base = buf.data();
        while (((size_t)base & 31) != 0) base++;
        std::memcpy(base, in, sizeof(gr_complex) * ((size_t)(n - 1) * decimate + ntaps));
        if (decimate == 1) f.filterN(reinterpret_cast<gr_complex *>(out), base, (unsigned long)n);
        else f.filterNdec(reinterpret_cast<gr_complex *>(out), base, (unsigned long)n, (unsigned)decimate);
        return 0;
    } catch (...) {
        return -1;
    }
}

/* fft_complex(fft_size, forward).execute() (lib/fft.cc:146-242): the plan wrapper clFFT_impl::testCPU uses */
REF_API int ref_fft_complex(int n, int forward, const float *in, float *out, long nvec)
{
    try {
        fft_complex f(n, forward != 0, 1);
        for (long v = 0; v < nvec; v++) {
            std::memcpy(f.get_inbuf(), in + 2 * (size_t)v * n, sizeof(gr_complex) * n);
            f.execute();
            std::memcpy(out + 2 * (size_t)v * n, f.get_outbuf(), sizeof(gr_complex) * n);
        }
        return 0;
    } catch (...) {
        return -1;
    }
}
