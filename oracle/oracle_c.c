/*
 * oracle/oracle_c.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the arithmetic of the gr-clenabled streaming DSP hot
 * path.  It is the checker the CUDA kernels are compared
 * against and the "port" CPU baseline that bench.py times; nothing under
 * gr_clenabled_b200/ may link, import or call it.
 *
 * PARITY PINNING.  Pinned against the reference's OWN code compiled here from
 * /root/reference (oracle/Makefile -> oracle/_ref/, vectors under tests/golden/):
 *   - window / firdes tap design: lib/window.cc + lib/firdes.cc
 *     (libref_firdes.so, ref_firdes_window.npz);
 *   - fft_filter_ccf, fir_filter_ccf and the fft_complex plan wrapper:
 *     lib/fft_filter.cc + lib/fir_filter.cc + lib/fft.cc against the stand-in
 *     VOLK / FFTW3 / Boost headers of oracle/shim/ (libref_filters.so,
 *     ref_filters.npz): blocking, overlap-add tail, tap reversal, decimation
 *     counters are the reference's code; the innermost library kernels (FFT
 *     butterflies, dot products) are the shim's;
 *   - numeric ids: include/clenabled/*.h + grc/*.yml (ref_constants.json).
 *   - the device arithmetic itself: the OpenCL kernel strings XCorrelate, CharToComplex (IChar and packed-XY LUT),
 *     filterpfb2 + channel_map and opconst_complex, emitted by the reference's OWN builder functions (their bodies
 *     are compiled from where they lie into a generator) and run on the CPU against oracle/shim/opencl_c.h
 *     (oracle/ref_kernels.py, ref_kernels.npz).
 * Not runnable here: clFFT (third-party library: the DFT it is planned for is taken from its published definition and
 * checked against pocketfft) and the reference's host classes around the kernels (GNU Radio, OpenCL C++ API).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference tree).
 *
 * Build: make -C oracle   ->  oracle/liboracle.so   (gcc -O3 -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* op codes: include/clenabled/clMathOpTypes.h:11-20 */
enum {
    OP_MULTIPLY = 1, OP_ADD = 2, OP_SUBTRACT = 3, OP_CONJ = 4, OP_MULCONJ = 5,
    OP_EMPTY = 255, OP_EMPTY_W_COPY = 254
};

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* synthetic input RNG: SURVEY.md 8(d) -- splitmix64(seed + index)           */
/* ------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* top 24 bits -> uniform float in [-1,1) */
ORC_API void orc_rng_f32(float *out, long n, uint64_t seed, uint64_t first)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        uint64_t r = splitmix64(seed + first + (uint64_t)i);
        out[i] = (float)((double)(r >> 40) / 8388608.0 - 1.0);
    }
}

/* top 8 bits as int8, -128 mapped to -127 (symmetric range) */
ORC_API void orc_rng_i8(int8_t *out, long n, uint64_t seed, uint64_t first)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        uint64_t r = splitmix64(seed + first + (uint64_t)i);
        int v = (int)(int8_t)(r >> 56);
        out[i] = (int8_t)(v == -128 ? -127 : v);
    }
}

/* ------------------------------------------------------------------------ */
/* M1/M2: clMathConst  (lib/clMathConst_impl.cc:169-222 kernel,              */
/*        :275-301 testCPU).  k is a REAL float applied to re and im.        */
/* ------------------------------------------------------------------------ */
ORC_API void orc_mathconst_c32(const float *in, float *out, long nitems, float k, int op)
{
    long n = 2 * nitems;
    switch (op) {
    case OP_EMPTY_W_COPY:   /* testCPU :278-283 copies; the OpenCL kernel string falls
                             * through into the multiply (missing break, :187-193) --
                             * the CPU behaviour is the one restated */
        memcpy(out, in, sizeof(float) * n);
        break;
    case OP_MULTIPLY:
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) out[i] = in[i] * k;
        break;
    case OP_ADD:
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) out[i] = in[i] + k;
        break;
    case OP_SUBTRACT:
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) out[i] = in[i] - k;
        break;
    case OP_CONJ:           /* :203-218  c.imag = -1.0 * a.imag */
#pragma omp parallel for schedule(static)
        for (long i = 0; i < nitems; i++) {
            out[2 * i] = in[2 * i];
            out[2 * i + 1] = -in[2 * i + 1];
        }
        break;
    default:                /* MATHOP_EMPTY: kernel returns, output untouched */
        break;
    }
}

/* opconst_float (:121-143) */
ORC_API void orc_mathconst_f32(const float *in, float *out, long n, float k, int op)
{
    for (long i = 0; i < n; i++) {
        if (op == OP_MULTIPLY) out[i] = in[i] * k;
        else if (op == OP_ADD) out[i] = in[i] + k;
        else if (op == OP_SUBTRACT) out[i] = in[i] - k;
    }
}

/* opconst_int (:144-167): multiplier is (int)k */
ORC_API void orc_mathconst_i32(const int32_t *in, int32_t *out, long n, float k, int op)
{
    int32_t m = (int32_t)k;
    for (long i = 0; i < n; i++) {
        /* wrap-around like the device: do the arithmetic unsigned */
        if (op == OP_MULTIPLY) out[i] = (int32_t)((uint32_t)in[i] * (uint32_t)m);
        else if (op == OP_ADD) out[i] = (int32_t)((uint32_t)in[i] + (uint32_t)m);
        else if (op == OP_SUBTRACT) out[i] = (int32_t)((uint32_t)in[i] - (uint32_t)m);
    }
}

/* ------------------------------------------------------------------------ */
/* M4: clMathOp op_complex (lib/clMathOp_impl.cc:178-236), testCPU :336-352  */
/* Each product/sum is a separately rounded float op (no FMA contraction);   */
/* compiled with -ffp-contract=off.                                          */
/* ------------------------------------------------------------------------ */
ORC_API void orc_mathop_c32(const float *a, const float *b, float *c, long nitems, int op)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < nitems; i++) {
        float ar = a[2 * i], ai = a[2 * i + 1];
        float br = b[2 * i], bi = b[2 * i + 1];
        switch (op) {
        case OP_MULTIPLY:
            c[2 * i] = (ar * br) - (ai * bi);
            c[2 * i + 1] = (ar * bi) + (ai * br);
            break;
        case OP_MULCONJ:    /* b_i = -1.0 * b.imag  (:228) */
            bi = -bi;
            c[2 * i] = (ar * br) - (ai * bi);
            c[2 * i + 1] = (ar * bi) + (ai * br);
            break;
        case OP_ADD:
            c[2 * i] = ar + br;
            c[2 * i + 1] = ai + bi;
            break;
        case OP_SUBTRACT:
            c[2 * i] = ar - br;
            c[2 * i + 1] = ai - bi;
            break;
        default:
            break;
        }
    }
}

/* op_float (:120-147) */
ORC_API void orc_mathop_f32(const float *a, const float *b, float *c, long n, int op)
{
    for (long i = 0; i < n; i++) {
        if (op == OP_MULTIPLY) c[i] = a[i] * b[i];
        else if (op == OP_ADD) c[i] = a[i] + b[i];
        else if (op == OP_SUBTRACT) c[i] = a[i] - b[i];
    }
}

/* op_int (:150-175; the reference source has a typo and never compiles, the
 * intended semantics are restated) */
ORC_API void orc_mathop_i32(const int32_t *a, const int32_t *b, int32_t *c, long n, int op)
{
    for (long i = 0; i < n; i++) {
        if (op == OP_MULTIPLY) c[i] = (int32_t)((uint32_t)a[i] * (uint32_t)b[i]);
        else if (op == OP_ADD) c[i] = (int32_t)((uint32_t)a[i] + (uint32_t)b[i]);
        else if (op == OP_SUBTRACT) c[i] = (int32_t)((uint32_t)a[i] - (uint32_t)b[i]);
    }
}

/* ------------------------------------------------------------------------ */
/* M5: secondary element-wise kernels, evaluated in double then rounded      */
/* (these are tolerance-compared, 1e-5 relative)                             */
/* ------------------------------------------------------------------------ */
/* clLog: lib/clLog_impl.cc:55-57,139-148  c = (n/log2(10))*log2(a) + k      */
ORC_API void orc_log10(const float *a, float *c, long n, float nval, float kval)
{
    double f = (double)nval / log2(10.0);
    for (long i = 0; i < n; i++) c[i] = (float)(f * log2((double)a[i]) + (double)kval);
}

/* clSNR: lib/clSNR_impl.cc:105-113  c = fabs(n*log10(a/b)+k) */
ORC_API void orc_snr(const float *a, const float *b, float *c, long n, float nval, float kval)
{
    for (long i = 0; i < n; i++) {
        float t = a[i] / b[i];
        c[i] = (float)fabs((double)nval * log10((double)t) + (double)kval);
    }
}

/* clComplexToMag: lib/clComplexToMag_impl.cc:140-148 */
ORC_API void orc_complex_to_mag(const float *a, float *c, long n)
{
    for (long i = 0; i < n; i++) {
        double re = a[2 * i], im = a[2 * i + 1];
        c[i] = (float)sqrt(im * im + re * re);
    }
}

/* clComplexToArg: lib/clComplexToArg_impl.cc:139-151 (double atan2 branch) */
ORC_API void orc_complex_to_arg(const float *a, float *c, long n)
{
    for (long i = 0; i < n; i++) c[i] = (float)atan2((double)a[2 * i + 1], (double)a[2 * i]);
}

/* clComplexToMagPhase: lib/clComplexToMagPhase_impl.cc:151-165 */
ORC_API void orc_complex_to_magphase(const float *a, float *mag, float *ph, long n)
{
    orc_complex_to_mag(a, mag, n);
    orc_complex_to_arg(a, ph, n);
}

/* clMagPhaseToComplex: lib/clMagPhaseToComplex_impl.cc:169-192 (double branch) */
ORC_API void orc_magphase_to_complex(const float *mag, const float *ph, float *c, long n)
{
    for (long i = 0; i < n; i++) {
        c[2 * i] = (float)((double)mag[i] * cos((double)ph[i]));
        c[2 * i + 1] = (float)((double)mag[i] * sin((double)ph[i]));
    }
}

/* ------------------------------------------------------------------------ */
/* window / firdes restatement (lib/window.cc:95-104,139-170;                */
/* lib/firdes.cc:93-135,675-686) -- pinned against oracle/_ref               */
/* ------------------------------------------------------------------------ */
ORC_API void orc_window_blackman(float *w, int ntaps)
{
    float M = (float)(ntaps - 1);
    const float c0 = 0.42f, c1 = 0.5f, c2 = 0.08f;
    for (int n = 0; n < ntaps; n++)
        w[n] = c0 - c1 * cosf((2.0f * M_PI * n) / M) + c2 * cosf((4.0f * M_PI * n) / M);
}

ORC_API void orc_window_hamming(float *w, int ntaps)
{
    float M = (float)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) w[n] = 0.54 - 0.46 * cos((2 * M_PI * n) / M);
}

/* firdes::compute_ntaps with WIN_HAMMING (max attenuation 53 dB) */
ORC_API int orc_firdes_ntaps_hamming(double fs, double tw)
{
    int ntaps = (int)(53.0 * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    return ntaps;
}

/* firdes::low_pass with a Hamming window; taps must hold ntaps (odd) floats */
ORC_API void orc_firdes_low_pass_hamming(float *taps, int ntaps, double gain, double fs, double fc)
{
    float *w = (float *)malloc(sizeof(float) * ntaps);
    orc_window_hamming(w, ntaps);
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * M_PI * fc / fs;
    for (int n = -M; n <= M; n++) {
        if (n == 0) taps[n + M] = fwT0 / M_PI * w[n + M];
        else taps[n + M] = sin(n * fwT0) / (n * M_PI) * w[n + M];
    }
    double fmax = taps[0 + M];
    for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M];
    gain /= fmax;
    for (int i = 0; i < ntaps; i++) taps[i] *= gain;
    free(w);
}

/* ------------------------------------------------------------------------ */
/* F1/F4: clFFT.  Unnormalised DFT, float32 arithmetic, twiddles computed in */
/* double and rounded (clFFT_impl.cc:104-106 comment).  Stands in for FFTW3f */
/* (lib/fft.cc:175-179) / clFFT, neither of which is available here.         */
/* Iterative radix-2 decimation-in-time on a bit-reversed copy.              */
/* ------------------------------------------------------------------------ */
typedef struct {
    int n, logn;
    float *tw;          /* n/2 complex twiddles e^{-2 pi i k/n} */
    uint32_t *rev;
} orc_fft_plan;

static orc_fft_plan *fft_plan_make(int n)
{
    int logn = 0;
    while ((1 << logn) < n) logn++;
    if ((1 << logn) != n) return NULL;
    orc_fft_plan *p = (orc_fft_plan *)malloc(sizeof(*p));
    p->n = n;
    p->logn = logn;
    p->tw = (float *)malloc(sizeof(float) * (n > 1 ? n : 2));
    p->rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (int k = 0; k < n / 2; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        p->tw[2 * k] = (float)cos(a);
        p->tw[2 * k + 1] = (float)sin(a);
    }
    for (int i = 0; i < n; i++) {
        uint32_t r = 0;
        for (int b = 0; b < logn; b++)
            if (i & (1 << b)) r |= 1u << (logn - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

static void fft_plan_free(orc_fft_plan *p)
{
    if (!p) return;
    free(p->tw);
    free(p->rev);
    free(p);
}

/* plans are built once per size and kept, like the reference's fft_complex plan that lives as long as its block
 * (lib/fft.cc:146-192): a timed per-call loop must not pay for twiddle generation */
static orc_fft_plan *fft_plan_cached(int n)
{
    static orc_fft_plan *cache[32];
    int logn = 0;
    while ((1 << logn) < n) logn++;
    if ((1 << logn) != n || logn >= 32) return NULL;
    orc_fft_plan *p;
#pragma omp critical(orc_plan_cache)
    {
        if (!cache[logn]) cache[logn] = fft_plan_make(n);
        p = cache[logn];
    }
    return p;
}

/* in-place on x (interleaved), x already in bit-reversed order; sign=-1 fwd */
static void fft_exec_bitrev(const orc_fft_plan *p, float *x, int sign)
{
    int n = p->n;
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < half; j++) {
                float wr = p->tw[2 * j * step];
                float wi = p->tw[2 * j * step + 1];
                if (sign > 0) wi = -wi;
                float *a = x + 2 * (i + j), *b = x + 2 * (i + j + half);
                float tr = b[0] * wr - b[1] * wi;
                float ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] += tr;
                a[1] += ti;
            }
        }
    }
}

static void fft_exec(const orc_fft_plan *p, const float *in, float *out, int sign)
{
    int n = p->n;
    for (int i = 0; i < n; i++) {
        uint32_t r = p->rev[i];
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    fft_exec_bitrev(p, out, sign);
}

/* Sizes that are not a power of two (clFFT plans accept 2^a 3^b 5^c 7^d, lib/clFFT_impl.cc:97-100): the transform is
 * the DFT definition itself, X[k] = sum_n x[n] e^{sign 2 pi i n k / N}, accumulated in double -- O(N^2), test sizes only. */
static void dft_direct(const float *in, float *out, int n, int sign)
{
    double *cs = (double *)malloc(sizeof(double) * 2 * n);
    for (int i = 0; i < n; i++) {
        double a = sign * 2.0 * M_PI * (double)i / (double)n;
        cs[2 * i] = cos(a);
        cs[2 * i + 1] = sin(a);
    }
    for (int k = 0; k < n; k++) {
        double re = 0, im = 0;
        for (int i = 0; i < n; i++) {
            int idx = (int)(((long)i * k) % n);
            re += in[2 * i] * cs[2 * idx] - in[2 * i + 1] * cs[2 * idx + 1];
            im += in[2 * i] * cs[2 * idx + 1] + in[2 * i + 1] * cs[2 * idx];
        }
        out[2 * k] = (float)re;
        out[2 * k + 1] = (float)im;
    }
    free(cs);
}

static void fft_any(const orc_fft_plan *p, const float *in, float *out, int n, int sign)
{
    if (p) fft_exec(p, in, out, sign);
    else dft_direct(in, out, n, sign);
}

/*
 * clFFT_impl::processOpenCL (lib/clFFT_impl.cc:526-634), complex input.
 *  dir: -1 forward (CLFFT_FORWARD, e^{-i..}), +1 backward; scale 1.0 both ways (:121-122)
 *  backward && shift: the two halves of the INPUT are swapped on upload (:548-553)
 *  window (may be NULL): a[i] *= w[i] on the uploaded (already swapped) buffer (:566-580)
 *  forward && shift: the two halves of the OUTPUT are swapped afterwards (:594-607)
 */
ORC_API int orc_fft_c32(const float *in, float *out, int n, long nvec, int dir,
                        const float *window, int shift)
{
    orc_fft_plan *p = fft_plan_cached(n);      /* NULL: not a power of two -> direct DFT */
    if (n < 1) return -1;
    int fwd = (dir < 0);
    int h = n / 2;                             /* vlen_2 (:81): an odd size leaves its last element where it is */
#pragma omp parallel
    {
        float *a = (float *)malloc(sizeof(float) * 2 * n);
        float *c = (float *)malloc(sizeof(float) * 2 * n);
#pragma omp for schedule(static)
        for (long v = 0; v < nvec; v++) {
            const float *x = in + 2 * (size_t)n * v;
            float *y = out + 2 * (size_t)n * v;
            memcpy(a, x, sizeof(float) * 2 * n);
            if (!fwd && shift) {
                memcpy(a, x + 2 * h, sizeof(float) * 2 * h);
                memcpy(a + 2 * h, x, sizeof(float) * 2 * h);
            }
            if (window) {
                for (int i = 0; i < n; i++) {
                    a[2 * i] *= window[i];
                    a[2 * i + 1] *= window[i];
                }
            }
            fft_any(p, a, c, n, fwd ? -1 : +1);
            memcpy(y, c, sizeof(float) * 2 * n);
            if (fwd && shift) {
                memcpy(y, c + 2 * h, sizeof(float) * 2 * h);
                memcpy(y + 2 * h, c, sizeof(float) * 2 * h);
            }
        }
        free(a);
        free(c);
    }
    return 0;
}

/*
 * Real-input forward path (lib/clFFT_impl.cc:556-565,608-630): R2C Hermitian
 * output then the upper half filled by conjugate symmetry.  The reference's
 * fill loop is off by one; the mathematically correct
 * full spectrum X[N-k] = conj(X[k]) is restated here.
 */
ORC_API int orc_fft_r32(const float *in, float *out, int n, long nvec, const float *window)
{
    orc_fft_plan *p = fft_plan_cached(n);
    if (n < 1) return -1;
    float *a = (float *)malloc(sizeof(float) * 2 * n);
    for (long v = 0; v < nvec; v++) {
        for (int i = 0; i < n; i++) {
            a[2 * i] = in[(size_t)n * v + i] * (window ? window[i] : 1.0f);
            a[2 * i + 1] = 0.0f;
        }
        fft_any(p, a, out + 2 * (size_t)n * v, n, -1);
    }
    free(a);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* L2: clFilter time domain.  td_FIR_complex (lib/clFilter_impl.cc:162-194): */
/*   out[g] = sum_{i<K} taps[K-1-i] * in[g+i], float accumulate in order i,  */
/* then the host-side decimation stride copy (:563-586): every D-th output.  */
/* `in` holds nin + K - 1 samples (history K-1 in front, :78).               */
/* Returns the number of outputs written (= ceil(nin/D)).                    */
/* ------------------------------------------------------------------------ */
ORC_API long orc_fir_ccf(const float *in, float *out, long nin, const float *taps, int K, int D)
{
    long nout = (nin + D - 1) / D;
#pragma omp parallel for schedule(static)
    for (long o = 0; o < nout; o++) {
        long g = o * D;
        float re = 0.0f, im = 0.0f;
        for (int i = 0; i < K; i++) {
            float t = taps[K - 1 - i];
            re += t * in[2 * (g + i)];
            im += t * in[2 * (g + i) + 1];
        }
        out[2 * o] = re;
        out[2 * o + 1] = im;
    }
    return nout;
}

/* ------------------------------------------------------------------------ */
/* L1/L4: FFT filter, overlap-add (lib/fft_filter.cc:38-97 set_taps /        */
/* compute_sizes, :133-175 filter; GPU twin lib/clFilter_impl.cc:592-681).   */
/* ------------------------------------------------------------------------ */
typedef struct {
    int ntaps, fftsize, nsamples, decim;
    orc_fft_plan *plan;
    float *xtaps;       /* fftsize complex: FFT(taps/fftsize) */
    float *tail;        /* ntaps-1 complex */
    float *a, *b, *c;   /* work buffers */
} orc_fftfilt;

ORC_API void orc_fftfilt_sizes(int ntaps, int *fftsize, int *nsamples)
{
    /* fft_filter.cc:77-78 */
    int fs = (int)(2 * pow(2.0, ceil(log((double)ntaps) / log(2.0))));
    *fftsize = fs;
    *nsamples = fs - ntaps + 1;
}

ORC_API orc_fftfilt *orc_fftfilt_create(const float *taps, int ntaps, int decim)
{
    orc_fftfilt *f = (orc_fftfilt *)calloc(1, sizeof(*f));
    f->ntaps = ntaps;
    f->decim = decim;
    orc_fftfilt_sizes(ntaps, &f->fftsize, &f->nsamples);
    f->plan = fft_plan_make(f->fftsize);
    size_t nb = sizeof(float) * 2 * f->fftsize;
    f->xtaps = (float *)malloc(nb);
    f->a = (float *)malloc(nb);
    f->b = (float *)malloc(nb);
    f->c = (float *)malloc(nb);
    f->tail = (float *)calloc(2 * (ntaps > 1 ? ntaps - 1 : 1), sizeof(float));
    float scale = 1.0 / f->fftsize;                 /* fft_filter.cc:52 */
    memset(f->a, 0, nb);
    for (int i = 0; i < ntaps; i++) f->a[2 * i] = taps[i] * scale;
    fft_exec(f->plan, f->a, f->xtaps, -1);
    return f;
}

ORC_API void orc_fftfilt_destroy(orc_fftfilt *f)
{
    if (!f) return;
    fft_plan_free(f->plan);
    free(f->xtaps); free(f->tail); free(f->a); free(f->b); free(f->c);
    free(f);
}

/*
 * fft_filter_ccf::filter (fft_filter.cc:133-175).  nitems = number of OUTPUT
 * items; consumes nitems*decim inputs in blocks of nsamples.  Like GNU Radio's
 * own block (which sets output_multiple(nsamples)) the caller must pass
 * nitems*decim as a multiple of nsamples; dec_ctr restarts at 0 every call.
 */
ORC_API int orc_fftfilt_filter(orc_fftfilt *f, long nitems, const float *input, float *output)
{
    int dec_ctr = 0, j;
    long ninput = nitems * f->decim;
    int ns = f->nsamples, N = f->fftsize, tail = f->ntaps - 1;
    for (long i = 0; i < ninput; i += ns) {
        memcpy(f->a, input + 2 * i, sizeof(float) * 2 * ns);
        memset(f->a + 2 * ns, 0, sizeof(float) * 2 * (N - ns));
        fft_exec(f->plan, f->a, f->b, -1);
        for (int k = 0; k < N; k++) {           /* c = a*b (complex), :152 */
            float ar = f->b[2 * k], ai = f->b[2 * k + 1];
            float br = f->xtaps[2 * k], bi = f->xtaps[2 * k + 1];
            f->a[2 * k] = ar * br - ai * bi;
            f->a[2 * k + 1] = ar * bi + ai * br;
        }
        fft_exec(f->plan, f->a, f->c, +1);
        for (j = 0; j < tail; j++) {
            f->c[2 * j] += f->tail[2 * j];
            f->c[2 * j + 1] += f->tail[2 * j + 1];
        }
        j = dec_ctr;
        while (j < ns) {
            *output++ = f->c[2 * j];
            *output++ = f->c[2 * j + 1];
            j += f->decim;
        }
        dec_ctr = j - ns;
        memcpy(f->tail, f->c + 2 * ns, sizeof(float) * 2 * tail);
    }
    return (int)nitems;
}

/* ------------------------------------------------------------------------ */
/* P1: clPolyphaseChannelizer (lib/clPolyphaseChannelizer_impl.cc:156-177    */
/* kernels, :208-225 plan, :83-109 general_work).                            */
/*  in : buf_items + T - M samples, in[0] is T-1 samples in the past          */
/*  filt[i*M + (j + i*(M-R)) % M] = sum_{k=j,j+M,..<T} in[i*R - k + T-1]*taps[k] */
/*  fft[i*M + c] = sum_n filt[i*M+n] e^{+2 pi i n c / M}   (BACKWARD, scale 1) */
/*  out[i*nmap + j] = fft[i*M + map[j]]                                       */
/* The arm sum uses fma() like the kernel (:163).  When R < M the reference   */
/* reads past its uploaded buffer; callers here must supply                   */
/* (buf_items/R - 1)*R + T samples.                                           */
/* ------------------------------------------------------------------------ */
ORC_API int orc_pfb(const float *in, float *out, const float *taps, int T, int M, int R,
                    const int *map, int nmap, long niter)
{
    orc_fft_plan *p = fft_plan_make(M);
    if (!p) return -1;
#pragma omp parallel
    {
        float *filt = (float *)malloc(sizeof(float) * 2 * M);
        float *spec = (float *)malloc(sizeof(float) * 2 * M);
#pragma omp for schedule(static)
        for (long i = 0; i < niter; i++) {
            for (int j = 0; j < M; j++) {
                float re = 0.0f, im = 0.0f;
                for (int k = j; k < T; k += M) {
                    const float *x = in + 2 * (i * R - k + T - 1);
                    re = fmaf(x[0], taps[k], re);
                    im = fmaf(x[1], taps[k], im);
                }
                long slot = ((long)j + i * (long)(M - R)) % M;
                filt[2 * slot] = re;
                filt[2 * slot + 1] = im;
            }
            fft_exec(p, filt, spec, +1);
            for (int j = 0; j < nmap; j++) {
                out[2 * (i * nmap + j)] = spec[2 * map[j]];
                out[2 * (i * nmap + j) + 1] = spec[2 * map[j] + 1];
            }
        }
        free(filt);
        free(spec);
    }
    fft_plan_free(p);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* X2/X3: clXEngine.                                                         */
/*  input layout [t][station][chan][pol] (lib/clXEngine_impl.cc:767-768,     */
/*  host marshal :987-1058); IChar adds an innermost [re,im] byte pair.      */
/*  baseline k <-> (s1 >= s2): k = s1(s1+1)/2 + s2  (:744-750)                */
/*  V = sum_t x[s1] * conj(x[s2]):  re = ar*br + ai*bi, im = ai*br - ar*bi    */
/*  (:729-736); npol=2 emits XX,XY,YX,YY = (row X,col X),(row X,col Y),      */
/*  (row Y,col X),(row Y,col Y) at 4*i..4*i+3 (:776-806), i = f*nbl + k.      */
/* ------------------------------------------------------------------------ */

/* exact integer accumulators: out[(f*nbl + k)*npol*npol + p][2] int32 */
ORC_API void orc_xengine_i8_exact(const int8_t *in, int32_t *out, int A, int F, int T, int npol)
{
    int nbl = A * (A + 1) / 2;
    long frame = (long)A * F * npol;        /* complex items per time step */
#pragma omp parallel for schedule(static) collapse(2)
    for (int f = 0; f < F; f++) {
        for (int k = 0; k < nbl; k++) {
            int s1 = (int)(-0.5 + sqrt(0.25 + 2.0 * k));
            while ((s1 + 1) * (s1 + 2) / 2 <= k) s1++;     /* guard sqrt rounding */
            while (s1 * (s1 + 1) / 2 > k) s1--;
            int s2 = k - (s1 + 1) * s1 / 2;
            int64_t acc[4][2] = {{0}};
            for (int t = 0; t < T; t++) {
                const int8_t *r = in + 2 * (t * frame + ((long)s1 * F + f) * npol);
                const int8_t *c = in + 2 * (t * frame + ((long)s2 * F + f) * npol);
                for (int p1 = 0; p1 < npol; p1++)
                    for (int p2 = 0; p2 < npol; p2++) {
                        int ar = r[2 * p1], ai = r[2 * p1 + 1];
                        int br = c[2 * p2], bi = c[2 * p2 + 1];
                        acc[p1 * npol + p2][0] += ar * br + ai * bi;
                        acc[p1 * npol + p2][1] += ai * br - ar * bi;
                    }
            }
            long i = (long)f * nbl + k;
            for (int p = 0; p < npol * npol; p++) {
                out[2 * (i * npol * npol + p)] = (int32_t)acc[p][0];
                out[2 * (i * npol * npol + p) + 1] = (int32_t)acc[p][1];
            }
        }
    }
}

/*
 * Reference-order float emulation: CharToComplex (:861-866) scales each byte
 * by the DOUBLE literal 1/127 and rounds to float, then XCorrelate (:739-810)
 * accumulates float products in t order (non-FMA branch :731-735).
 * Also serves complex-float input when in_f32 != NULL.
 */
ORC_API void orc_xengine_f32(const int8_t *in_i8, const float *in_f32, float *out,
                             int A, int F, int T, int npol, int accumulate)
{
    int nbl = A * (A + 1) / 2;
    long frame = (long)A * F * npol;
    const double s = 0.007874015748031496063;
#pragma omp parallel for schedule(static) collapse(2)
    for (int f = 0; f < F; f++) {
        for (int k = 0; k < nbl; k++) {
            int s1 = (int)(-0.5 + sqrt(0.25 + 2.0 * k));
            while ((s1 + 1) * (s1 + 2) / 2 <= k) s1++;
            while (s1 * (s1 + 1) / 2 > k) s1--;
            int s2 = k - (s1 + 1) * s1 / 2;
            float acc[4][2] = {{0}};
            for (int t = 0; t < T; t++) {
                long i1 = t * frame + ((long)s1 * F + f) * npol;
                long i2 = t * frame + ((long)s2 * F + f) * npol;
                for (int p1 = 0; p1 < npol; p1++)
                    for (int p2 = 0; p2 < npol; p2++) {
                        float ar, ai, br, bi;
                        if (in_f32) {
                            ar = in_f32[2 * (i1 + p1)]; ai = in_f32[2 * (i1 + p1) + 1];
                            br = in_f32[2 * (i2 + p2)]; bi = in_f32[2 * (i2 + p2) + 1];
                        } else {
                            ar = (float)((float)in_i8[2 * (i1 + p1)] * s);
                            ai = (float)((float)in_i8[2 * (i1 + p1) + 1] * s);
                            br = (float)((float)in_i8[2 * (i2 + p2)] * s);
                            bi = (float)((float)in_i8[2 * (i2 + p2) + 1] * s);
                        }
                        acc[p1 * npol + p2][0] += ar * br + ai * bi;
                        acc[p1 * npol + p2][1] += ai * br - ar * bi;
                    }
            }
            long i = (long)f * nbl + k;
            for (int p = 0; p < npol * npol; p++) {
                float *o = out + 2 * (i * npol * npol + p);
                if (accumulate) { o[0] += acc[p][0]; o[1] += acc[p][1]; }
                else { o[0] = acc[p][0]; o[1] = acc[p][1]; }
            }
        }
    }
}

/* Packed 4-bit LUT of CharToComplex (:833): {0..7, 0, -7..-1}; hi nibble first */
ORC_API void orc_unpack4(const uint8_t *in, int8_t *out, long nbytes)
{
    static const int8_t lut[16] = {0, 1, 2, 3, 4, 5, 6, 7, 0, -7, -6, -5, -4, -3, -2, -1};
    for (long i = 0; i < nbytes; i++) {
        out[2 * i] = lut[in[i] >> 4];
        out[2 * i + 1] = lut[in[i] & 0x0F];
    }
}

/* ======================================================================================
 * SURVEY 8(f) "next" rows: the reference correlators, clComplexFilter, clQuadratureDemod,
 * clSignalSource.
 * ====================================================================================== */

/* clXCorrelate ComplexToMag kernel (lib/clXCorrelate_impl.cc:915-929): float fma form */
ORC_API void orc_xc_mag(const float *a, float *c, long n)
{
    for (long i = 0; i < n; i++) c[i] = sqrtf(fmaf(a[2 * i], a[2 * i], a[2 * i + 1] * a[2 * i + 1]));
}

/* max_shift as the constructor derives it (lib/clXCorrelate_impl.cc:727-747): the search
 * index (or 0.7 * signal_length made even) rounded UP to a power of two */
ORC_API int orc_xc_max_shift(int signal_length, int max_search_index)
{
    int ms;
    if (max_search_index > 0) {
        ms = max_search_index;
    } else {
        ms = (int)(0.7f * (float)signal_length);
        if (ms % 2) ms += 1;
    }
    float p2 = log2f((float)ms);
    int nms = (int)pow(2.0, ceil((double)p2));
    return nms;
}

/* kernel XCorrelate (lib/clXCorrelate_impl.cc:851-900): one normalised correlation factor
 * per shift g - max_shift, g in [0, 2*max_shift); xx/yy are the squared magnitudes
 * (F32Squared, :976-980); sequential float sums; -2.0 where the overlap has no energy */
ORC_API void orc_xc_factors(const float *ref, const float *sig, int L, int max_shift, float *factors)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int g = 0; g < 2 * max_shift; g++) {
        int shift = g - max_shift;
        int ref_start = shift >= 0 ? shift : -shift;
        int calc_len = L - ref_start;
        float sum_xy = 0, sum_x2 = 0, sum_y2 = 0;
        if (shift > 0) {
            for (int i = 0; i < calc_len; i++) {
                sum_xy += ref[ref_start + i] * sig[i];
                sum_x2 += ref[ref_start + i] * ref[ref_start + i];
                sum_y2 += sig[i] * sig[i];
            }
        } else {
            for (int i = 0; i < calc_len; i++) {
                sum_xy += ref[i] * sig[ref_start + i];
                sum_x2 += ref[i] * ref[i];
                sum_y2 += sig[ref_start + i] * sig[ref_start + i];
            }
        }
        float denom = sum_x2 * sum_y2;
        factors[g] = (denom != 0.0f) ? sum_xy / sqrtf(sum_x2 * sum_y2) : -2.0f;
    }
}

/* find_max (kernel :1016-1043 + host pass :1371-1413): work-groups of `group` consecutive
 * entries reduced by a halving tree that replaces slot id by slot id+stride only when
 * STRICTLY greater, then the first strictly greatest group wins.  Ties therefore go to the
 * lower SLOT at every level, which is not always the lowest index. */
ORC_API void orc_xc_find_max(const float *in, int n, int group, float *corr, int *index)
{
    int ngroups = n / group;
    float best = 0;
    int best_i = 0;
    float *m = (float *)malloc(sizeof(float) * group);
    int *l = (int *)malloc(sizeof(int) * group);
    for (int gi = 0; gi < ngroups; gi++) {
        for (int k = 0; k < group; k++) {
            m[k] = in[gi * group + k];
            l[k] = gi * group + k;
        }
        for (int stride = group / 2; stride > 0; stride /= 2)
            for (int k = 0; k < stride; k++)
                if (m[k + stride] > m[k]) {
                    m[k] = m[k + stride];
                    l[k] = l[k + stride];
                }
        if (gi == 0 || m[0] > best) {
            best = m[0];
            best_i = l[0];
        }
    }
    free(m);
    free(l);
    *corr = best;
    *index = best_i;
}

/* clxcorrelate_fft_vcf::work (lib/clxcorrelate_fft_vcf_impl.cc:1057-1145): per vector,
 * [forward FFT of both when input_type == 2,] ref * conj(sig) (MultConj :895-906), backward FFT
 * with scale 1.0 (:727), magnitude (:925-931), halves swapped (:1136-1141) */
ORC_API int orc_xcorr_fft_vcf(const float *ref, const float *sig, float *out, int n, long nvec, int input_type)
{
    orc_fft_plan *p = fft_plan_make(n);
    if (!p) return -1;
    int h = n / 2;
#pragma omp parallel
    {
        float *a = (float *)malloc(sizeof(float) * 2 * n);
        float *b = (float *)malloc(sizeof(float) * 2 * n);
        float *c = (float *)malloc(sizeof(float) * 2 * n);
#pragma omp for schedule(static)
        for (long v = 0; v < nvec; v++) {
            const float *r = ref + 2 * (size_t)n * v, *s = sig + 2 * (size_t)n * v;
            if (input_type == 2) {
                memcpy(c, r, sizeof(float) * 2 * n);
                fft_exec(p, c, a, -1);
                memcpy(c, s, sizeof(float) * 2 * n);
                fft_exec(p, c, b, -1);
            } else {
                memcpy(a, r, sizeof(float) * 2 * n);
                memcpy(b, s, sizeof(float) * 2 * n);
            }
            for (int i = 0; i < n; i++) {
                float a_r = a[2 * i], a_i = a[2 * i + 1], b_r = b[2 * i], b_i = -b[2 * i + 1];
                c[2 * i] = (a_r * b_r) - (a_i * b_i);
                c[2 * i + 1] = (a_r * b_i) + (a_i * b_r);
            }
            fft_exec(p, c, a, +1);
            float *y = out + (size_t)n * v;
            for (int i = 0; i < n; i++) {
                float m = sqrtf(fmaf(a[2 * i], a[2 * i], a[2 * i + 1] * a[2 * i + 1]));
                y[(i + h) % n] = m;
            }
        }
        free(a);
        free(b);
        free(c);
    }
    fft_plan_free(p);
    return 0;
}

/* clComplexFilter td_FIR_complex_complex (lib/clComplexFilter_impl.cc:805-829): complex taps,
 * out[g] = sum_{i<K} taps[K-1-i] * in[g+i], in[0] is K-1 samples old (set_history); decimation
 * keeps every D-th output as the host loop of clFilter does.  Returns outputs written. */
ORC_API long orc_fir_ccc(const float *in, float *out, long nin, const float *taps, int K, int D)
{
    long nfull = nin - (K - 1);
    long nout = 0;
    for (long g = 0; g < nfull; g += D, nout++) {
        float re = 0.0f, im = 0.0f;
        for (int i = 0; i < K; i++) {
            float a_r = taps[2 * (K - 1 - i)], a_i = taps[2 * (K - 1 - i) + 1];
            float b_r = in[2 * (g + i)], b_i = in[2 * (g + i) + 1];
            re += (a_r * b_r) - (a_i * b_i);
            im += (a_r * b_i) + (a_i * b_r);
        }
        out[2 * nout] = re;
        out[2 * nout + 1] = im;
    }
    return nout;
}

/* clQuadratureDemod quadDemod, double + fma branch (lib/clQuadratureDemod_impl.cc:126-143):
 * out[i] = gain * atan2(Im, Re) of in[i+1] * conj(in[i]); in holds n+1 samples (set_history(2)) */
ORC_API void orc_quad_demod(const float *in, float *out, long n, float gain)
{
    for (long i = 0; i < n; i++) {
        double a_r = in[2 * (i + 1)], a_i = in[2 * (i + 1) + 1];
        double b_r = in[2 * i], b_i = -1.0 * in[2 * i + 1];
        double re = fma(a_r, b_r, -(a_i * b_i));
        double im = fma(a_r, b_i, a_i * b_r);
        out[i] = (gain != 1.0f) ? (float)((double)gain * atan2(im, re)) : (float)atan2(im, re);
    }
}

/* clSignalSource sig_float / sig_complex, double branch (lib/clSignalSource_impl.cc:128-211):
 * dval = phase + phase_inc * index; waveform 0 = cos, 1 = sin (float) or (cos, sin) (complex) */
ORC_API void orc_sig_source(float *out, long n, int complex_out, int waveform_sin, double phase, double phase_inc,
                            double ampl)
{
    for (long i = 0; i < n; i++) {
        double d = phase + phase_inc * (double)i;
        if (complex_out) {
            out[2 * i] = (float)(cos(d) * ampl);
            out[2 * i + 1] = (float)(sin(d) * ampl);
        } else {
            out[i] = (float)((waveform_sin ? sin(d) : cos(d)) * ampl);
        }
    }
}

/* phase bookkeeping after a call (lib/clSignalSource_impl.cc:386-398): advance by
 * inc * (float)n and wrap into (-2pi, 2pi) by dropping whole turns */
ORC_API double orc_sig_source_advance(double phase, double phase_inc, long n)
{
    const double two_pi = 6.28318530717958647692;
    phase = phase + (phase_inc * (float)n);
    if (phase > two_pi || phase < -two_pi) {
        phase = phase / two_pi - (double)((int)(phase / two_pi));
        phase = phase * two_pi;
    }
    return phase;
}
