/* oracle/shim/volk/volk.h -- TEST INFRASTRUCTURE ONLY.
 * The handful of VOLK entry points the reference's lib/fft.cc, lib/fft_filter.cc and lib/fir_filter.cc
 * call, as plain sequential loops (what VOLK's generic kernels compute; the SIMD kernels differ from
 * them only in float summation order).  VOLK itself is not in the image.  Lets oracle/Makefile compile
 * those reference files from where they lie. */
#pragma once
/* the real headers pull these in transitively; the reference relies on it (std::reverse, memcpy, pow) */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

typedef std::complex<float> lv_32fc_t;

static inline size_t volk_get_alignment(void) { return 32; }
static inline void *volk_malloc(size_t size, size_t alignment)
{
    void *p = nullptr;
    if (posix_memalign(&p, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : 1) != 0) return nullptr;
    return p;
}
static inline void volk_free(void *p) { free(p); }

static inline void volk_32fc_x2_multiply_32fc_a(lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) {
        const float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
        c[i] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
    }
}
static inline void volk_32fc_conjugate_32fc(lv_32fc_t *c, const lv_32fc_t *a, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) c[i] = std::conj(a[i]);
}
static inline void volk_32f_x2_dot_prod_32f_a(float *r, const float *a, const float *b, unsigned int n)
{
    float s = 0.f;
    for (unsigned int i = 0; i < n; i++) s += a[i] * b[i];
    *r = s;
}
static inline void volk_32fc_32f_dot_prod_32fc_a(lv_32fc_t *r, const lv_32fc_t *a, const float *b, unsigned int n)
{
    float sr = 0.f, si = 0.f;
    for (unsigned int i = 0; i < n; i++) {
        sr += a[i].real() * b[i];
        si += a[i].imag() * b[i];
    }
    *r = lv_32fc_t(sr, si);
}
static inline void volk_32fc_x2_dot_prod_32fc_a(lv_32fc_t *r, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int n)
{
    float sr = 0.f, si = 0.f;
    for (unsigned int i = 0; i < n; i++) {
        sr += a[i].real() * b[i].real() - a[i].imag() * b[i].imag();
        si += a[i].real() * b[i].imag() + a[i].imag() * b[i].real();
    }
    *r = lv_32fc_t(sr, si);
}
static inline void volk_16i_32fc_dot_prod_32fc_a(lv_32fc_t *r, const short *a, const lv_32fc_t *b, unsigned int n)
{
    float sr = 0.f, si = 0.f;
    for (unsigned int i = 0; i < n; i++) {
        sr += (float)a[i] * b[i].real();
        si += (float)a[i] * b[i].imag();
    }
    *r = lv_32fc_t(sr, si);
}
static inline void volk_32f_x2_dot_prod_16i_a(int16_t *r, const float *a, const float *b, unsigned int n)
{
    float s = 0.f;
    for (unsigned int i = 0; i < n; i++) s += a[i] * b[i];
    *r = (int16_t)s;
}
