// fft.cu -- clFFT block: batched 1-D complex FFT (window, shift fused), sm_100a.
//
// Replaces clFFT_impl::processOpenCL (lib/clFFT_impl.cc:526-634): where the
// reference issues, per vector, an H2D copy, an optional MultiplyFloat window
// kernel (:202-229), one clfftEnqueueTransform with batch 1 (:583), a D2H copy
// and then a host memcpy fftshift (:594-607), this does every vector of a
// work() call in ONE launch: the window multiply is fused into the first
// pass' HBM load, the backward-shift into its load index (:548-553) and the
// forward-shift into the last pass' store index.
//
// HBM traffic: 8 B read + 8 B written per sample = 16 B/sample (algorithmic
// minimum); the transform itself lives in shared memory/registers.
#include "common.cuh"
#ifndef CLB_TW_DERIVE
#define CLB_TW_DERIVE 2      // see fft_device.cuh
#endif
#include "fft_device.cuh"
#include <cmath>
#include <cstdlib>

using namespace clb200;
using namespace clb200::fftdev;

namespace {


// MODE bit 0: inverse (re/im swapped on load and store), bit 1: real input, bit 2: pass-1 twiddles in shared memory,
// bit 3: tiles from the work counter (common.cuh: tile_fetch) instead of static striding
template <int LOGN, int EPT, int BATCH, int MINB, int MODE>
__global__ void __launch_bounds__((1 << LOGN) / EPT * BATCH, MINB)
k_fft(const float2 *__restrict__ in, float2 *__restrict__ out, long nvec,
      const float2 *__restrict__ tw, const float *__restrict__ win, int shift, unsigned long long *wq)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, T = P::T;
    constexpr bool inverse = MODE & 1, real_in = MODE & 2, tw1_smem = (MODE & 4) && P::npass() > 1;
    extern __shared__ __align__(16) float2 smem[];
    const float2 *tw1s = nullptr;
    if constexpr (tw1_smem) {
        constexpr int N1 = P::tw_offset(2 < P::npass() ? 2 : P::npass()) - P::tw_offset(1);
        float2 *dst = smem + BATCH * P::SMEM_F2;
        for (int i = threadIdx.x; i < N1; i += T * BATCH) dst[i] = __ldg(tw + P::tw_offset(1) + i);
        __syncthreads();
        tw1s = dst;
    }

    const int tb = (BATCH == 1) ? 0 : threadIdx.x / T;     // transform within the CTA
    const int lt = (BATCH == 1) ? threadIdx.x : threadIdx.x % T;
    float2 *buf = smem + tb * P::SMEM_F2;
    // forward+shift swaps the OUTPUT halves, backward+shift swaps the INPUT halves.
    // (lt + c) ^ N/2 == lt + (c ^ N/2) for the per-thread offsets c (multiples of T),
    // so the half swap is a choice between two base pointers per compile-time c.
    const int in_x = (shift && inverse) ? (N >> 1) : 0;
    const int out_x = (shift && !inverse) ? (N >> 1) : 0;
    // 2-, 4- and 8-point transforms (one thread each, one pass): vector I/O when both pointers are 16 B aligned
    constexpr bool one_thread = T == 1 && !real_in && N >= 2 && P::npass() == 1 && EPT == N;
    const bool vec_io = one_thread && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;

    const long ntile = (nvec + BATCH - 1) / BATCH;
    auto process = [&](long tile) {
        const long v = tile * BATCH + tb;
        const bool active = v < nvec;
        float2 x[EPT];

        if (active) {
            const float2 *src_up = in + v * N + lt + in_x, *src_dn = in + v * N + lt - in_x;
            const float *srf_up = reinterpret_cast<const float *>(in) + v * N + lt + in_x;
            const float *srf_dn = reinterpret_cast<const float *>(in) + v * N + lt - in_x;
            if (vec_io) {
                // one thread per transform: its N samples are contiguous -- 128-bit loads (two samples), through L1
                // so that the sectors a warp's strided requests share are fetched once
                if constexpr (one_thread && N == 2) {         // the half swap exchanges the two samples
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(in + v * N));
                    const float2 s0 = in_x ? make_float2(a.z, a.w) : make_float2(a.x, a.y);
                    const float2 s1 = in_x ? make_float2(a.x, a.y) : make_float2(a.z, a.w);
                    x[0] = inverse ? make_float2(s0.y, s0.x) : s0;
                    x[1] = inverse ? make_float2(s1.y, s1.x) : s1;
                } else if constexpr (one_thread) {
                    static_for<0, EPT / 2>([&](auto q_) {
                        constexpr int c = 2 * decltype(q_)::value;
                        const float4 a = __ldg(reinterpret_cast<const float4 *>(((c & (N >> 1)) ? src_dn : src_up) + c));
                        x[c] = inverse ? make_float2(a.y, a.x) : make_float2(a.x, a.y);
                        x[c + 1] = inverse ? make_float2(a.w, a.z) : make_float2(a.z, a.w);
                    });
                }
            } else
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                constexpr int c = in_index<P, EPT>(0, e);
                float2 a;
                if constexpr (real_in) a = make_float2(__ldcs(((c & (N >> 1)) ? srf_dn : srf_up) + c), 0.0f);
                else a = ldg_stream(((c & (N >> 1)) ? src_dn : src_up) + c);
                x[e] = inverse ? make_float2(a.y, a.x) : a;
            });
            if (win != nullptr) {
                const float *wp = win + lt;
                static_for<0, EPT>([&](auto e_) {
                    constexpr int e = decltype(e_)::value;
                    const float w = __ldg(wp + in_index<P, EPT>(0, e));
                    x[e].x *= w;
                    x[e].y *= w;
                });
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
        }

        fft_core<P, EPT>(x, buf, lt, tw, NoHook{}, tw1s);

        if (active) {
            float2 *dst_up = out + v * N + lt + out_x, *dst_dn = out + v * N + lt - out_x;
            if (vec_io) {
                if constexpr (one_thread && N == 2) {
                    const float2 a = out_x ? x[1] : x[0], b = out_x ? x[0] : x[1];
                    const float4 o = inverse ? make_float4(a.y, a.x, b.y, b.x) : make_float4(a.x, a.y, b.x, b.y);
                    __stcs(reinterpret_cast<float4 *>(out + v * N), o);
                } else if constexpr (one_thread) {
                    constexpr int LN = ilog2(N);                 // single pass: X[k] = x[bitrev(k)]
                    static_for<0, EPT / 2>([&](auto q_) {
                        constexpr int c = 2 * decltype(q_)::value;
                        const float2 a = x[bitrev(c, LN)], b = x[bitrev(c + 1, LN)];
                        const float4 o = inverse ? make_float4(a.y, a.x, b.y, b.x) : make_float4(a.x, a.y, b.x, b.y);
                        __stcs(reinterpret_cast<float4 *>(((c & (N >> 1)) ? dst_dn : dst_up) + c), o);
                    });
                }
            } else
            for_each_output_c<P, EPT>(x, [&](auto c_, float2 a) {
                constexpr int c = decltype(c_)::value;
                if (inverse) a = make_float2(a.y, a.x);
                __stcs(((c & (N >> 1)) ? dst_dn : dst_up) + c, a);
            });
        }
    };
    if constexpr ((MODE & 16) && P::npass() > 1) {
        // same hand-over, the loop written with the fetch as a run-time choice: ptxas schedules the tile body
        // differently around it, and which form is faster depends on the size (tools/fft_dyn_ab.py; pick_variant)
        __shared__ long s_next;
        const bool dyn = wq != nullptr;
        for (long tile = blockIdx.x; tile < ntile;) {
            long nxt = tile + gridDim.x;
            if (dyn && threadIdx.x == 0) nxt = tile_fetch(wq);
            process(tile);
            if (dyn) {
                if (threadIdx.x == 0) s_next = nxt;
                __syncthreads();
                nxt = s_next;
            }
            tile = nxt;
        }
        if (dyn && threadIdx.x == 0) tile_finish(wq);
    } else if constexpr ((MODE & 8) && P::npass() > 1) {
        // the barriers inside fft_core order the hand-over through s_next
        __shared__ long s_next;
        for (long tile = blockIdx.x; tile < ntile;) {
            long nxt = 0;
            if (threadIdx.x == 0) nxt = tile_fetch(wq);
            process(tile);
            if (threadIdx.x == 0) s_next = nxt;
            __syncthreads();
            tile = s_next;
        }
        if (threadIdx.x == 0) tile_finish(wq);
    } else {
        for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) process(tile);
    }
}

// clxcorrelate_fft_vcf (lib/clxcorrelate_fft_vcf_impl.cc:1057-1145) in ONE launch per signal
// pair instead of the reference's write / MultConj / backward clFFT / ComplexToMag / read /
// host memcpy-shift per vector: the product ref * conj(sig) (:895-906) is formed while the
// first pass loads the two spectra, the backward transform (scale 1.0, :727) runs in shared
// memory, and magnitude (:925-931) + the half swap (:1136-1141) are fused into the last
// pass' store.  24 B/sample of HBM traffic (2 x 8 in, 4 out... plus 4 B padding-free float out).
template <int LOGN, int EPT, int BATCH, int MINB>
__global__ void __launch_bounds__((1 << LOGN) / EPT * BATCH, MINB)
k_xcfft(const float2 *__restrict__ ref, const float2 *__restrict__ sig, float *__restrict__ out, long nvec,
        const float2 *__restrict__ tw)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, T = P::T;
    extern __shared__ __align__(16) float2 smem[];
    const int tb = (BATCH == 1) ? 0 : threadIdx.x / T;
    const int lt = (BATCH == 1) ? threadIdx.x : threadIdx.x % T;
    float2 *buf = smem + tb * P::SMEM_F2;
    const long ntile = (nvec + BATCH - 1) / BATCH;
    for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long v = tile * BATCH + tb;
        const bool active = v < nvec;
        float2 x[EPT];
        if (active) {
            const float2 *ra = ref + v * N + lt, *sa = sig + v * N + lt;
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                constexpr int c = in_index<P, EPT>(0, e);
                const float2 a = ldg_stream(ra + c), b = ldg_stream(sa + c);
                const float b_i = -b.y;
                const float re = (a.x * b.x) - (a.y * b_i), im = (a.x * b_i) + (a.y * b.x);
                x[e] = make_float2(im, re);                  // backward = forward on swapped re/im
            });
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
        }
        fft_core<P, EPT>(x, buf, lt, tw);
        if (active) {
            float *dst_up = out + v * N + lt + (N >> 1), *dst_dn = out + v * N + lt - (N >> 1);
            for_each_output_c<P, EPT>(x, [&](auto c_, float2 a) {
                constexpr int c = decltype(c_)::value;
                // (re, im) of the backward transform = (a.y, a.x)
                __stcs(((c & (N >> 1)) ? dst_dn : dst_up) + c, sqrtf(fmaf(a.y, a.y, a.x * a.x)));
            });
        }
    }
}

// Prefetching variant for the large sizes (one transform per CTA, complex input):
// the next vector is pulled into the (padded) working buffer by the bulk-copy engine
// -- one 16*G-byte row per thread, completion on an mbarrier -- as soon as the
// last pass has read its inputs, so the HBM latency of vector i+1 hides behind the
// last butterflies and the output stores of vector i.
template <int LOGN, int EPT, int MINB, int MODE>
__global__ void __launch_bounds__((1 << LOGN) / EPT, MINB)
k_fft_pf(const float2 *__restrict__ in, float2 *__restrict__ out, long nvec,
         const float2 *__restrict__ tw, const float *__restrict__ win, int shift)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, T = P::T, G = P::PADG, NROWS = N / G;
    constexpr bool inverse = MODE & 1;
    static_assert(T % G == 0 && P::npass() > 1, "prefetch kernel is for the multi-pass sizes");
    extern __shared__ __align__(16) float2 smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P::SMEM_F2);
    const int lt = threadIdx.x;
    const int in_x = (shift && inverse) ? (N >> 1) : 0;
    const int out_x = (shift && !inverse) ? (N >> 1) : 0;

    if (lt == 0) mbar_init(bar, 1);
    __syncthreads();
    auto issue = [&](long v) {
        if (lt == 0) mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
#pragma unroll
        for (int row = lt; row < NROWS; row += T)
            bulk_g2s(smem + P::pad(row * G), in + v * N + ((row * G) ^ in_x), G * (uint32_t)sizeof(float2), bar);
    };
    long v = blockIdx.x;
    uint32_t parity = 0;
    if (v < nvec) issue(v);
    const float2 *const ldp = smem + P::pad(lt);
    for (; v < nvec; v += gridDim.x) {
        mbar_wait(bar, parity);
        parity ^= 1;
        float2 x[EPT];
        static_for<0, EPT>([&](auto e_) {
            constexpr int e = decltype(e_)::value;
            const float2 a = ldp[P::pad(in_index<P, EPT>(0, e))];
            x[e] = inverse ? make_float2(a.y, a.x) : a;
        });
        if (win != nullptr) {
            const float *wp = win + lt;
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                const float w = __ldg(wp + in_index<P, EPT>(0, e));
                x[e].x *= w;
                x[e].y *= w;
            });
        }
        const long vn = v + gridDim.x;
        fft_core<P, EPT>(x, smem, lt, tw, [&]() {
            __syncthreads();                    // every thread has its last-pass inputs
            if (vn < nvec) {
                fence_proxy_async();
                issue(vn);
            }
        });
        float2 *dst_up = out + v * N + lt + out_x, *dst_dn = out + v * N + lt - out_x;
        for_each_output_c<P, EPT>(x, [&](auto c_, float2 a) {
            constexpr int c = decltype(c_)::value;
            if (inverse) a = make_float2(a.y, a.x);
            __stcs(((c & (N >> 1)) ? dst_dn : dst_up) + c, a);
        });
    }
}

// The two transforms of the chirp-z path (lengths that are not a power of two, fft_launch_blue) with its element-wise
// steps fused in: STAGE 0 reads the caller's n-point vectors (window, real input, backward half swap, conjugation for
// the backward transform), multiplies by the chirp, zero-pads to the 2^LOGN-point block and transforms it; STAGE 1
// multiplies the block by the resident spectrum of the wrapped conj(chirp) on load, runs the inverse transform and
// stores the first n outputs times the chirp (forward half swap fused).  Four crossings of HBM instead of ten.
struct FftCz {
    const float2 *chirp;     // w[k], k < n
    const float2 *bspec;     // FFT_m of the wrapped conj(w), pre-scaled by 1/m
    const float *win;        // or null
    int n, real_in, inverse, swap_h;
};
__device__ __forceinline__ int cz_swap(int p, int h) { return p < h ? p + h : (p < 2 * h ? p - h : p); }

template <int LOGN, int EPT, int BATCH, int MINB, int STAGE>
__global__ void __launch_bounds__((1 << LOGN) / EPT * BATCH, MINB)
k_fft_cz(const void *__restrict__ in, float2 *__restrict__ out, long nvec, const float2 *__restrict__ tw, FftCz cz)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, T = P::T;
    extern __shared__ __align__(16) float2 smem[];
    const int tb = (BATCH == 1) ? 0 : threadIdx.x / T;
    const int lt = (BATCH == 1) ? threadIdx.x : threadIdx.x % T;
    float2 *buf = smem + tb * P::SMEM_F2;
    const int n = cz.n;
    const long ntile = (nvec + BATCH - 1) / BATCH;
    for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long v = tile * BATCH + tb;
        const bool active = v < nvec;
        float2 x[EPT];
        static_for<0, EPT>([&](auto e_) {
            constexpr int e = decltype(e_)::value;
            const int idx = in_index<P, EPT>(lt, e);
            float2 a = make_float2(0.f, 0.f);
            if (STAGE == 0) {
                if (active && idx < n) {
                    const int src = cz.swap_h ? cz_swap(idx, cz.swap_h) : idx;
                    a = cz.real_in ? make_float2(__ldcs(reinterpret_cast<const float *>(in) + v * n + src), 0.f)
                                   : ldg_stream(reinterpret_cast<const float2 *>(in) + v * n + src);
                    if (cz.win != nullptr) {
                        const float w = __ldg(cz.win + idx);
                        a.x *= w;
                        a.y *= w;
                    }
                    if (cz.inverse) a.y = -a.y;
                    a = cmul(a, __ldg(cz.chirp + idx));
                }
            } else {
                if (active) {
                    a = cmul(ldg_stream(reinterpret_cast<const float2 *>(in) + v * N + idx), __ldg(cz.bspec + idx));
                    a = make_float2(a.y, a.x);                          // inverse transform = forward on swapped re/im
                }
            }
            x[e] = a;
        });
        fft_core<P, EPT>(x, buf, lt, tw);
        if (active) {
            for_each_output<P, EPT>(x, lt, [&](int k, float2 a) {
                if (STAGE == 0) {
                    __stcs(out + v * N + k, a);
                } else if (k < n) {
                    a = cmul(make_float2(a.y, a.x), __ldg(cz.chirp + k));
                    if (cz.inverse) a.y = -a.y;
                    __stcs(out + v * n + (cz.swap_h ? cz_swap(k, cz.swap_h) : k), a);
                }
            });
        }
    }
}

// ------------------------------------------------------------------ host ----
struct FftVariant {
    int logn, ept, batch, threads, smem_bytes, tw_total, max_ctas_per_sm;
    void (*fill_tw)(std::vector<float2> &);
    void (*kernel[3])(const float2 *, float2 *, long, const float2 *, const float *, int, unsigned long long *);
    void (*kernel_pf[2])(const float2 *, float2 *, long, const float2 *, const float *, int);   // or null
    void (*kernel_xc)(const float2 *, const float2 *, float *, long, const float2 *);           // FFT correlator
    void (*kernel_s[2])(const float2 *, float2 *, long, const float2 *, const float *, int, unsigned long long *);    // pass-1 twiddles in smem
    void (*kernel_d[3])(const float2 *, float2 *, long, const float2 *, const float *, int, unsigned long long *);    // tiles from the work counter (multi-pass sizes) or null
    void (*kernel_u[3])(const float2 *, float2 *, long, const float2 *, const float *, int, unsigned long long *);    // the same, loop form 2
    int tw1_bytes;
    void (*kernel_cz[2])(const void *, float2 *, long, const float2 *, FftCz);    // chirp-z stages (fft_launch_blue)
};

template <int LOGN, int EPT>
void fill_tw_t(std::vector<float2> &tw)
{
    using P = Plan<LOGN, EPT>;
    tw.assign(std::max(1, P::TW_TOTAL), make_float2(1.f, 0.f));
    for (int p = 1; p < P::npass(); p++) {
        int R = P::radix(p), NS = P::ns(p), off = P::tw_offset(p);
        for (int r = 1; r < R; r++)
            for (int k = 0; k < NS; k++) {
                double a = -2.0 * M_PI * (double)r * (double)k / ((double)NS * (double)R);
                tw[off + (r - 1) * NS + k] = make_float2((float)cos(a), (float)sin(a));
            }
    }
}

// which work-counter loop form a size uses (0: static striding)
constexpr int fft_loop_form(int logn)
{
    return (logn >= 6 && logn <= 13) ? 1 : logn == 14 ? 2 : 0;
}

template <int LOGN, int EPT, int BATCH, int MINB>
FftVariant make_variant()
{
    using P = Plan<LOGN, EPT>;
    FftVariant v;
    v.logn = LOGN;
    v.ept = EPT;
    v.batch = BATCH;
    v.threads = P::T * BATCH;
    v.smem_bytes = P::SMEM_F2 * BATCH * (int)sizeof(float2);
    v.tw_total = P::TW_TOTAL;
    v.max_ctas_per_sm = MINB;
    v.fill_tw = &fill_tw_t<LOGN, EPT>;
    v.kernel[0] = &k_fft<LOGN, EPT, BATCH, MINB, 0>;     // forward
    v.kernel[1] = &k_fft<LOGN, EPT, BATCH, MINB, 1>;     // backward
    v.kernel[2] = &k_fft<LOGN, EPT, BATCH, MINB, 2>;     // forward, real input
    v.kernel_pf[0] = v.kernel_pf[1] = nullptr;
    v.kernel_xc = &k_xcfft<LOGN, EPT, BATCH, MINB>;
    v.kernel_cz[0] = &k_fft_cz<LOGN, EPT, BATCH, MINB, 0>;
    v.kernel_cz[1] = &k_fft_cz<LOGN, EPT, BATCH, MINB, 1>;
    v.kernel_d[0] = v.kernel_d[1] = v.kernel_d[2] = nullptr;
    // measured per size on B200 (tools/fft_dyn_ab.py -> profiles/r2_tile_ab.txt), round-1 instantiation with static
    // striding = 100 %: with tiles of >= 2048 samples, the fewest passes and -- where the transform allows 128-thread
    // CTAs -- four CTAs per SM (64: 8x8 x 32 transforms per CTA, 128: 16x8, 512: 32x16, 1024: 32x32, 2048: 32x32x2 with
    // two transforms per 128-thread CTA, 4096: 32x32x4 with one) and work-counter tiles -- 64 points 116 %, 128: 111 %,
    // 256: 114 %, 512: 108 %, 1024: 119 %, 2048: 120 %, 4096: 115 %, 8192: 110 %, 16384: 105 % (loop form 2); 16 and 32
    // points keep their small static tiles (larger or dynamic ones measured slower)
    if constexpr (P::npass() > 1 && fft_loop_form(LOGN) == 1) {
        v.kernel_d[0] = &k_fft<LOGN, EPT, BATCH, MINB, 8>;
        v.kernel_d[1] = &k_fft<LOGN, EPT, BATCH, MINB, 9>;
        v.kernel_d[2] = &k_fft<LOGN, EPT, BATCH, MINB, 10>;
    }
    v.kernel_u[0] = v.kernel_u[1] = v.kernel_u[2] = nullptr;
    if constexpr (P::npass() > 1 && fft_loop_form(LOGN) == 2) {
        v.kernel_u[0] = &k_fft<LOGN, EPT, BATCH, MINB, 16>;
        v.kernel_u[1] = &k_fft<LOGN, EPT, BATCH, MINB, 17>;
        v.kernel_u[2] = &k_fft<LOGN, EPT, BATCH, MINB, 18>;
    }
    v.kernel_s[0] = &k_fft<LOGN, EPT, BATCH, MINB, 4>;
    v.kernel_s[1] = &k_fft<LOGN, EPT, BATCH, MINB, 5>;
    v.tw1_bytes = P::npass() > 1 ? (P::tw_offset(2 < P::npass() ? 2 : P::npass()) - P::tw_offset(1)) * (int)sizeof(float2) : 0;
    if constexpr (BATCH == 1 && LOGN >= 12) {
        v.kernel_pf[0] = &k_fft_pf<LOGN, EPT, MINB, 0>;
        v.kernel_pf[1] = &k_fft_pf<LOGN, EPT, MINB, 1>;
    }
    return v;
}

// one instantiation per size, {LOGN, EPT, transforms per CTA, min CTAs/SM}.  (The 8192-point alternates the rounds
// measured -- 16 or 8 elements per thread, one CTA per SM: 2.9-4.0 TB/s, DESIGN.md 4.1 -- are no longer compiled in.)
const FftVariant *pick_variant(int logn)
{
    static const FftVariant tab[] = {
        make_variant<1, 2, 256, 8>(),   make_variant<2, 4, 32, 24>(),  make_variant<3, 8, 32, 16>(),
        make_variant<4, 4, 8, 24>(),    make_variant<5, 8, 8, 24>(),   make_variant<6, 8, 32, 2>(),
        make_variant<7, 16, 16, 4>(),   make_variant<8, 16, 16, 2>(),  make_variant<9, 32, 16, 2>(),
        make_variant<10, 32, 8, 2>(),   make_variant<11, 32, 2, 4>(),  make_variant<12, 32, 1, 4>(),
        make_variant<13, 32, 1, 2>(),   make_variant<14, 32, 1, 1>(),
    };
    if (logn < 1 || logn > 14) return nullptr;
    return &tab[logn - 1];
}

struct ColVariant;
const ColVariant *pick_col(int logn);
struct Fft;
int fft_setup_two_pass(Fft *f, const ColVariant *ca, const ColVariant *cb);

struct Fft : clb200_block {
    int n = 0, logn = 0, dir = 0, dtype = 0, shift = 0, mode = 0;
    bool has_window = false;
    const FftVariant *var = nullptr;
    Buf d_tw, d_win;
    int resident = 1;     // CTAs per SM the launch is sized for
    bool use_pf = false;  // bulk-copy prefetching kernel available and enabled
    bool use_tw1s = false; // pass-1 twiddles in shared memory (k_fft<.., MODE|4>)
    bool dyn_ok = false;   // work-counter kernel available with the same occupancy
    int loop_form = 1;     // which of the two work-counter loop forms (kernel_d / kernel_u)
    // sizes above 16384 (one CTA's shared memory): N = n1 x n2, two passes of the in-SM kernels around transposes
    bool big = false;
    // 32768 .. 1048576 points: two passes of column transforms (k_fft_col) instead of the five of the four-step path
    bool two_pass = false;
    const ColVariant *colA = nullptr, *colB = nullptr;
    Buf d_twA, d_twB, d_tw4;     // twiddles of the two column plans; W_N^j (j < 512) followed by W_N^(512 j) (j < N / 512)
    int col_resident[2] = {1, 1};
    // sizes that are not a power of two: chirp-z (Bluestein) on top of two power-of-two plans of m >= 2n-1 points
    bool blue = false, blue_fused = false;
    int m = 0, blue_resident = 1;
    Buf d_chirp, d_bspec;        // w[n] = e^{-i pi n^2 / N} (n floats2), FFT_m of the wrapped conj(w), pre-scaled by 1/m
    int n1 = 0, n2 = 0;
    clb200_handle sub1 = nullptr, sub2 = nullptr;      // n1- and n2-point plans (complex, no window, no shift)
    struct Scratch {
        cudaStream_t st = nullptr;
        Buf a, b;
    } scratch[4];                                      // per launching stream (the host path rotates over 3)
    ~Fft() override
    {
        DeviceGuard g(device);
        d_tw.release();
        d_win.release();
        d_chirp.release();
        d_bspec.release();
        d_twA.release();
        d_twB.release();
        d_tw4.release();
        for (auto &sc : scratch) {
            sc.a.release();
            sc.b.release();
        }
        if (sub1) clb200_destroy(sub1);
        if (sub2) clb200_destroy(sub2);
    }
};

// ---- sizes above 16384: four-step decomposition --------------------------------------------------------------
// x[n], n = n2 + N2*n1   ->   X[k1 + N1*k2] = sum_n2 [ W_N^(n2 k1) ( sum_n1 x[N2 n1 + n2] W_N1^(n1 k1) ) ] W_N2^(n2 k2)
// as: transpose-in (window, backward half swap, real -> complex)  ->  N1-point FFTs over contiguous vectors
//     ->  twiddle + transpose  ->  N2-point FFTs  ->  transpose-out (forward half swap).
// Every FFT pass is the in-SM kernel at its HBM rate; the three transposes are what the size costs.
// out[v][c][r] = f(in[v][r][c]), 32 x 32 tiles through shared memory, both sides coalesced
template <int MODE>      // 0: in, 1: twiddle, 2: out
__global__ void __launch_bounds__(256) k_fft_tr(const void *__restrict__ in, float2 *__restrict__ out, int rows, int cols,
                                                long nvec, const float *__restrict__ win, int n, int xor_idx, float sign,
                                                int real_in)
{
    __shared__ float2 tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (long v = blockIdx.z; v < nvec; v += gridDim.z) {
        const long base = v * (long)n;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            float2 val = make_float2(0.f, 0.f);
            if (r < rows && c < cols) {
                const int idx = r * cols + c;
                if (MODE == 0) {
                    const int src = idx ^ xor_idx;                       // backward + shift: input halves swapped
                    val = real_in ? make_float2(reinterpret_cast<const float *>(in)[base + src], 0.f)
                                  : reinterpret_cast<const float2 *>(in)[base + src];
                    if (win != nullptr) {                                // the window multiplies the (swapped) buffer
                        const float w = win[idx];
                        val.x *= w;
                        val.y *= w;
                    }
                } else {
                    val = reinterpret_cast<const float2 *>(in)[base + idx];
                    if (MODE == 1) {                                     // W_N^(n2 k1): row = n2, column = k1
                        const unsigned ph = ((unsigned)r * (unsigned)c) & (unsigned)(n - 1);
                        float sn, cs;
                        sincospif(2.0f * (float)ph / (float)n, &sn, &cs);
                        sn *= sign;
                        val = make_float2(val.x * cs - val.y * sn, val.x * sn + val.y * cs);
                    }
                }
            }
            tile[ty + 8 * k][tx] = val;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + ty + 8 * k, r = r0 + tx;                  // transposed: consecutive lanes -> consecutive r
            if (r < rows && c < cols) {
                int o = c * rows + r;
                if (MODE == 2) o ^= xor_idx;                             // forward + shift: output halves swapped
                out[base + o] = tile[tx][ty + 8 * k];
            }
        }
        __syncthreads();
    }
}

int fft_launch(Fft *f, const void *d_in, void *d_out, long nvec, cudaStream_t st);
Fft::Scratch *fft_scratch(Fft *f, cudaStream_t st, size_t bytes, int *rc);

// ---- 32768 .. 1048576 points in TWO passes over HBM: column transforms ------------------------------------------
// The four-step decomposition above spends three of its five passes on transposes.  Here a CTA owns C = 16 adjacent
// COLUMNS of the [NA][NB] view of a vector (128 B runs of every row, CTAs that run together take neighbouring runs of
// the same rows), thread = (column, element group) with the column fastest so that the global accesses coalesce; each
// column has its own shared-memory line of ODD length (16 columns x the same element = 16 distinct bank pairs).
//   pass A: x as [N1][N2], columns n2: window / backward half swap / real input fused into the load, N1-point
//           transforms, then through shared memory once more so that the store runs along k1:
//           T'[n2][k1] = W_N^(n2 k1) X_n2[k1]  (full rows, coalesced);
//   pass B: T' as [N2][N1], columns k1: N2-point transforms, stored straight from registers to X[k2 N1 + k1]
//           (forward half swap fused).
// The backward transform is the forward one on re/im-swapped data (swapped by pass A's load and pass B's store).
template <int LOGN, int EPT, int C, int MINB, int PASSB>
__global__ void __launch_bounds__((1 << LOGN) / EPT * C, MINB)
k_fft_col(const void *__restrict__ in, float2 *__restrict__ out, long nvec, int nb, int n, const float2 *__restrict__ tw,
          const float *__restrict__ win, int xor_idx, int real_in, int inverse, const float2 *__restrict__ tw4)
{
    using P = Plan<LOGN, EPT>;
    constexpr int NA = P::N, T = P::T, LINE = P::SMEM_F2 | 1;
    extern __shared__ __align__(16) float2 smem[];
    const int tb = threadIdx.x % C, lt = threadIdx.x / C;
    float2 *buf = smem + tb * LINE;
    const int groups = nb / C;
    const long ntile = nvec * groups;
    for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long v = tile / groups;
        const int c0 = (int)(tile - v * groups) * C;
        const long base = v * (long)n;
        float2 x[EPT];
        static_for<0, EPT>([&](auto e_) {
            constexpr int e = decltype(e_)::value;
            const int idx = in_index<P, EPT>(lt, e) * nb + c0 + tb;
            float2 a;
            if (PASSB) {
                a = reinterpret_cast<const float2 *>(in)[base + idx];
            } else {
                const int src = idx ^ xor_idx;                         // backward + shift: input halves swapped
                a = real_in ? make_float2(reinterpret_cast<const float *>(in)[base + src], 0.f)
                            : reinterpret_cast<const float2 *>(in)[base + src];
                if (win != nullptr) {
                    const float w = __ldg(win + idx);
                    a.x *= w;
                    a.y *= w;
                }
                if (inverse) a = make_float2(a.y, a.x);
            }
            x[e] = a;
        });
        fft_core<P, EPT, NoHook, 1, false>(x, buf, lt, tw);
        if (PASSB) {
            for_each_output<P, EPT>(x, lt, [&](int k, float2 a) {
                const int o = (k * nb + c0 + tb) ^ xor_idx;           // forward + shift: output halves swapped
                if (inverse) a = make_float2(a.y, a.x);
                __stcs(out + base + o, a);
            });
        } else {
            __syncthreads();                                           // the last pass has read its inputs
            for_each_output<P, EPT>(x, lt, [&](int k, float2 a) { buf[P::pad(k)] = a; });
            __syncthreads();
            // W_N^(n2 k1) = W_N^(512 hi) * W_N^lo from two L1-resident tables of <= 512 entries for every fourth column,
            // the three columns in between by multiplying with W_N^k1 (at most three extra roundings): the table reads
            // were as many LSU wavefronts as the data itself
            auto wn = [&](unsigned ph) {
                ph &= (unsigned)(n - 1);
                return cmul(__ldg(tw4 + 512 + (ph >> 9)), __ldg(tw4 + (ph & 511)));
            };
#pragma unroll
            for (int kk = 0; kk < NA / (T * C) + (NA % (T * C) != 0); kk++) {
                const int k = threadIdx.x + kk * T * C;
                if (k >= NA) break;
                const float2 step = wn((unsigned)k);
                float2 w = make_float2(1.f, 0.f);
#pragma unroll
                for (int col = 0; col < C; col++) {
                    if (col % 4 == 0) w = wn((unsigned)(c0 + col) * (unsigned)k);
                    else w = cmul(w, step);
                    const float2 a = cmul(smem[col * LINE + P::pad(k)], w);
                    __stcs(out + base + (long)(c0 + col) * NA + k, a);
                }
            }
        }
    }
}

typedef void (*fft_col_kernel_t)(const void *, float2 *, long, int, int, const float2 *, const float *, int, int, int, const float2 *);
struct ColVariant {
    int logn, threads, smem_bytes;
    void (*fill_tw)(std::vector<float2> &);
    fft_col_kernel_t kernel[2];      // pass A, pass B
};
template <int LOGN, int EPT, int MINB>
ColVariant make_col()
{
    using P = Plan<LOGN, EPT>;
    return ColVariant{LOGN, P::T * 16, (P::SMEM_F2 | 1) * 16 * (int)sizeof(float2), &fill_tw_t<LOGN, EPT>,
                      {&k_fft_col<LOGN, EPT, 16, MINB, 0>, &k_fft_col<LOGN, EPT, 16, MINB, 1>}};
}

const ColVariant *pick_col(int logn)
{
    static const ColVariant c7 = make_col<7, 16, 6>(), c8 = make_col<8, 16, 3>(), c9 = make_col<9, 32, 2>(),
                            c10 = make_col<10, 32, 1>();
    return logn == 7 ? &c7 : logn == 8 ? &c8 : logn == 9 ? &c9 : logn == 10 ? &c10 : nullptr;
}

int fft_launch_two_pass(Fft *f, const void *d_in, void *d_out, long nvec, cudaStream_t st)
{
    int rc = CLB200_OK;
    Fft::Scratch *sc = fft_scratch(f, st, (size_t)nvec * f->n * sizeof(float2), &rc);
    if (!sc) return rc;
    float2 *t = (float2 *)sc->a.p;
    const int N1 = f->n1, N2 = f->n2, half = f->n >> 1, sms = device_sm_count(f->device);
    const int inverse = f->dir > 0;
    const long tilesA = nvec * (N2 / 16), tilesB = nvec * (N1 / 16);
    f->colA->kernel[0]<<<grid_for(tilesA, sms, f->col_resident[0]), f->colA->threads, f->colA->smem_bytes, st>>>(
        d_in, t, nvec, N2, f->n, (const float2 *)f->d_twA.p, f->has_window ? (const float *)f->d_win.p : nullptr,
        (f->shift && inverse) ? half : 0, f->dtype == CLB200_DTYPE_FLOAT, inverse, (const float2 *)f->d_tw4.p);
    f->colB->kernel[1]<<<grid_for(tilesB, sms, f->col_resident[1]), f->colB->threads, f->colB->smem_bytes, st>>>(
        t, (float2 *)d_out, nvec, N1, f->n, (const float2 *)f->d_twB.p, nullptr, (f->shift && !inverse) ? half : 0, 0, inverse,
        nullptr);
    CLB_CUDA(cudaGetLastError());
    f->n_launch += 2;
    return CLB200_OK;
}

// ---- sizes that are not a power of two: Bluestein ------------------------------------------------------------
// clFFT plans accept any 2^a 3^b 5^c 7^d length (lib/clFFT_impl.cc:97-100 hands fftSize straight to
// clfftCreateDefaultPlan); here ANY length runs as a chirp-z transform over the power-of-two kernels:
//   n k = (n^2 + k^2 - (k - n)^2) / 2   =>   X[k] = w[k] * sum_n (x[n] w[n]) conj(w)[k - n],   w[n] = e^{-i pi n^2 / N}
// i.e. one circular convolution of length m >= 2N - 1 (m a power of two):  pre-multiply + zero-pad, FFT_m, multiply by
// the resident spectrum of the wrapped conj(w) (pre-scaled by 1/m), inverse FFT_m, post-multiply.  The backward
// transform is conj(DFT(conj x)).  Window, real input and the half swaps of the reference (vlen_2 = N / 2: an odd
// length leaves its last element in place, lib/clFFT_impl.cc:81,548-553,594-607) are fused into the two
// element-wise kernels.  Functional completeness (about ten passes over HBM), not a roofline kernel.
__device__ __forceinline__ int blue_swap(int p, int h)       // position p of a buffer whose halves [0,h) [h,2h) are swapped
{
    return p < h ? p + h : (p < 2 * h ? p - h : p);
}

__global__ void __launch_bounds__(256) k_blue_pre(const void *__restrict__ in, float2 *__restrict__ a, long nvec, int n, int m,
                                                  const float2 *__restrict__ chirp, const float *__restrict__ win, int real_in,
                                                  int inverse, int swap_h)
{
    const long total = nvec * (long)m;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long v = i / m;
        const int j = (int)(i - v * m);
        float2 val = make_float2(0.f, 0.f);
        if (j < n) {
            const int src = swap_h ? blue_swap(j, swap_h) : j;
            val = real_in ? make_float2(reinterpret_cast<const float *>(in)[v * n + src], 0.f)
                          : reinterpret_cast<const float2 *>(in)[v * n + src];
            if (win != nullptr) {
                const float w = win[j];
                val.x *= w;
                val.y *= w;
            }
            if (inverse) val.y = -val.y;
            val = cmul(val, chirp[j]);
        }
        a[i] = val;
    }
}

__global__ void __launch_bounds__(256) k_blue_mul(float2 *__restrict__ a, long total, int m, const float2 *__restrict__ bspec)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        a[i] = cmul(a[i], bspec[i & (m - 1)]);
}

__global__ void __launch_bounds__(256) k_blue_post(const float2 *__restrict__ c, float2 *__restrict__ out, long nvec, int n, int m,
                                                   const float2 *__restrict__ chirp, int inverse, int swap_h)
{
    const long total = nvec * (long)n;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long v = i / n;
        const int k = (int)(i - v * n);
        float2 val = cmul(c[v * m + k], chirp[k]);
        if (inverse) val.y = -val.y;
        out[v * n + (swap_h ? blue_swap(k, swap_h) : k)] = val;
    }
}

int fft_launch_blue(Fft *f, const void *d_in, void *d_out, long nvec, cudaStream_t st)
{
    int rc = CLB200_OK;
    Fft *s1 = static_cast<Fft *>(f->sub1);
    if (f->blue_fused && s1->var) {
        // m <= 16384: the two in-SM transforms with the element-wise steps fused in (k_fft_cz), one scratch buffer
        Fft::Scratch *sc = fft_scratch(f, st, (size_t)nvec * f->m * sizeof(float2), &rc);
        if (!sc) return rc;
        const FftVariant *v = s1->var;
        const long ntile = (nvec + v->batch - 1) / v->batch;
        const int grid = grid_for(ntile, device_sm_count(f->device), f->blue_resident);
        const int inverse = f->dir > 0, h = f->n / 2;
        FftCz cz{(const float2 *)f->d_chirp.p, (const float2 *)f->d_bspec.p, f->has_window ? (const float *)f->d_win.p : nullptr,
                 f->n, f->dtype == CLB200_DTYPE_FLOAT, inverse, (f->shift && inverse) ? h : 0};
        v->kernel_cz[0]<<<grid, v->threads, v->smem_bytes, st>>>(d_in, (float2 *)sc->a.p, nvec, (const float2 *)s1->d_tw.p, cz);
        cz.swap_h = (f->shift && !inverse) ? h : 0;
        v->kernel_cz[1]<<<grid, v->threads, v->smem_bytes, st>>>(sc->a.p, (float2 *)d_out, nvec, (const float2 *)s1->d_tw.p, cz);
        CLB_CUDA(cudaGetLastError());
        f->n_launch += 2;
        return CLB200_OK;
    }
    Fft::Scratch *sc = fft_scratch(f, st, (size_t)nvec * f->m * sizeof(float2), &rc);
    if (!sc) return rc;
    float2 *ta = (float2 *)sc->a.p, *tb = (float2 *)sc->b.p;
    const int N = f->n, M = f->m, h = N / 2;
    const int inverse = f->dir > 0;
    const long tot_m = nvec * (long)M, tot_n = nvec * (long)N;
    const int sms = device_sm_count(f->device);
    auto grid = [&](long total) { return (int)std::max<long>(1, std::min<long>((total + 255) / 256, (long)sms * 8)); };
    k_blue_pre<<<grid(tot_m), 256, 0, st>>>(d_in, ta, nvec, N, M, (const float2 *)f->d_chirp.p,
                                            f->has_window ? (const float *)f->d_win.p : nullptr, f->dtype == CLB200_DTYPE_FLOAT,
                                            inverse, (f->shift && inverse) ? h : 0);
    CLB_TRY(fft_launch(static_cast<Fft *>(f->sub1), ta, tb, nvec, st));
    k_blue_mul<<<grid(tot_m), 256, 0, st>>>(tb, tot_m, M, (const float2 *)f->d_bspec.p);
    CLB_TRY(fft_launch(static_cast<Fft *>(f->sub2), tb, ta, nvec, st));
    k_blue_post<<<grid(tot_n), 256, 0, st>>>(ta, (float2 *)d_out, nvec, N, M, (const float2 *)f->d_chirp.p, inverse,
                                             (f->shift && !inverse) ? h : 0);
    CLB_CUDA(cudaGetLastError());
    f->n_launch += 3;
    return CLB200_OK;
}

// the two scratch buffers of the launching stream, grown to `bytes` each
Fft::Scratch *fft_scratch(Fft *f, cudaStream_t st, size_t bytes, int *rc)
{
    Fft::Scratch *sc = nullptr;
    for (auto &c : f->scratch)
        if (c.st == st && c.a.p) sc = &c;
    if (!sc)
        for (auto &c : f->scratch)
            if (!c.a.p && !sc) sc = &c;
    if (!sc) {                                        // more launching streams than scratch sets: take over the first one
        sc = &f->scratch[0];
        if (cudaStreamSynchronize(sc->st) != cudaSuccess) {      // ... once its owner's work on it is done
            set_error("clFFT: scratch takeover failed");
            *rc = CLB200_ECUDA;
            return nullptr;
        }
    }
    sc->st = st;
    if ((*rc = sc->a.reserve(bytes)) != CLB200_OK || (*rc = sc->b.reserve(bytes)) != CLB200_OK) return nullptr;
    return sc;
}

int fft_launch_big(Fft *f, const void *d_in, void *d_out, long nvec, cudaStream_t st)
{
    int src = CLB200_OK;
    Fft::Scratch *sc = fft_scratch(f, st, (size_t)nvec * f->n * sizeof(float2), &src);
    if (!sc) return src;
    float2 *ta = (float2 *)sc->a.p, *tb = (float2 *)sc->b.p;
    const int N1 = f->n1, N2 = f->n2, half = f->n >> 1;
    const int gz = (int)std::min<long>(nvec, 64);
    const float sign = f->dir < 0 ? -1.f : 1.f;
    // x as [n1][n2] -> [n2][n1]
    k_fft_tr<0><<<dim3((N2 + 31) / 32, (N1 + 31) / 32, gz), 256, 0, st>>>(
        d_in, ta, N1, N2, nvec, f->has_window ? (const float *)f->d_win.p : nullptr, f->n,
        (f->shift && f->dir > 0) ? half : 0, sign, f->dtype == CLB200_DTYPE_FLOAT);
    CLB_TRY(fft_launch(static_cast<Fft *>(f->sub1), ta, tb, nvec * N2, st));            // over n1 -> [n2][k1]
    k_fft_tr<1><<<dim3((N1 + 31) / 32, (N2 + 31) / 32, gz), 256, 0, st>>>(tb, ta, N2, N1, nvec, nullptr, f->n, 0, sign, 0);
    CLB_TRY(fft_launch(static_cast<Fft *>(f->sub2), ta, tb, nvec * N1, st));            // over n2 -> [k1][k2]
    k_fft_tr<2><<<dim3((N2 + 31) / 32, (N1 + 31) / 32, gz), 256, 0, st>>>(tb, (float2 *)d_out, N1, N2, nvec, nullptr, f->n,
                                                                          (f->shift && f->dir < 0) ? half : 0, sign, 0);
    CLB_CUDA(cudaGetLastError());
    f->n_launch += 3;
    return CLB200_OK;
}

int fft_launch(Fft *f, const void *d_in, void *d_out, long nvec, cudaStream_t st)
{
    if (nvec <= 0) return CLB200_OK;
    if (f->blue) return fft_launch_blue(f, d_in, d_out, nvec, st);
    if (f->two_pass) return fft_launch_two_pass(f, d_in, d_out, nvec, st);
    if (f->big) return fft_launch_big(f, d_in, d_out, nvec, st);
    const FftVariant *v = f->var;
    long ntile = (nvec + v->batch - 1) / v->batch;
    int grid = grid_for(ntile, device_sm_count(f->device), f->resident);
    // tiles come from the work counter for the sizes that gain from it (fft_loop_form)
    unsigned long long *wq = (f->dyn_ok && ntile > grid) ? f->work_counter(st) : nullptr;
    // the bulk-copy kernel needs 16 B aligned rows; anything else takes the plain one
    const bool pf = f->use_pf && (((uintptr_t)d_in & 15) == 0);
    if (pf)
        v->kernel_pf[f->mode]<<<std::min<long>(grid, nvec), v->threads, v->smem_bytes + 16, st>>>(
            (const float2 *)d_in, (float2 *)d_out, nvec, (const float2 *)f->d_tw.p,
            f->has_window ? (const float *)f->d_win.p : nullptr, f->shift);
    else if (f->use_tw1s)
        v->kernel_s[f->mode]<<<grid, v->threads, v->smem_bytes + v->tw1_bytes, st>>>(
            (const float2 *)d_in, (float2 *)d_out, nvec, (const float2 *)f->d_tw.p,
            f->has_window ? (const float *)f->d_win.p : nullptr, f->shift, nullptr);
    else
        (wq ? (f->loop_form == 2 ? v->kernel_u[f->mode] : v->kernel_d[f->mode]) : v->kernel[f->mode])<<<grid, v->threads, v->smem_bytes, st>>>(
            (const float2 *)d_in, (float2 *)d_out, nvec, (const float2 *)f->d_tw.p,
            f->has_window ? (const float *)f->d_win.p : nullptr, f->shift, wq);
    CLB_CUDA(cudaGetLastError());
    f->n_launch++;
    return CLB200_OK;
}

// ------------------------------------------------------------- clxcorrelate_fft_vcf --
struct XcFft : clb200_block {
    int n = 0, num_inputs = 0, input_type = 1;
    const FftVariant *var = nullptr;
    Buf d_tw, d_spec;              // twiddles; forward spectra of every input (time-series mode)
    Buf d_in, d_out;               // host-path staging: [num_inputs][nvec][n] c32, [num_inputs-1][nvec][n] f32
    int resident = 1;
    cudaStream_t st = nullptr;
    ~XcFft() override
    {
        DeviceGuard g(device);
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        d_tw.release();
        d_spec.release();
        d_in.release();
        d_out.release();
    }
};

// d_in[k] : nvec vectors of input k (c32); d_out[k-1] : nvec vectors of float
int xcfft_launch(XcFft *x, const void *const *d_in, void *const *d_out, long nvec, cudaStream_t st)
{
    if (nvec <= 0) return CLB200_OK;
    const FftVariant *v = x->var;
    const long ntile = (nvec + v->batch - 1) / v->batch;
    const int grid = grid_for(ntile, device_sm_count(x->device), x->resident);
    const float2 *spec[CLB200_XCFFT_MAX_INPUTS];
    if (x->input_type == 2) {
        // time series: forward transform of every input first (:1079-1081, :1094-1096)
        const size_t per = (size_t)nvec * x->n * sizeof(float2);
        CLB_TRY(x->d_spec.reserve(per * x->num_inputs));
        for (int k = 0; k < x->num_inputs; k++) {
            float2 *dst = (float2 *)((char *)x->d_spec.p + per * k);
            v->kernel[0]<<<grid, v->threads, v->smem_bytes, st>>>((const float2 *)d_in[k], dst, nvec,
                                                                  (const float2 *)x->d_tw.p, nullptr, 0, nullptr);
            x->n_launch++;
            spec[k] = dst;
        }
    } else {
        for (int k = 0; k < x->num_inputs; k++) spec[k] = (const float2 *)d_in[k];
    }
    for (int k = 1; k < x->num_inputs; k++) {
        v->kernel_xc<<<grid, v->threads, v->smem_bytes, st>>>(spec[0], spec[k], (float *)d_out[k - 1], nvec,
                                                              (const float2 *)x->d_tw.p);
        x->n_launch++;
    }
    CLB_CUDA(cudaGetLastError());
    return CLB200_OK;
}

int fft_setup_two_pass(Fft *f, const ColVariant *ca, const ColVariant *cb)
{
    const ColVariant *cv[2] = {ca, cb};
    Buf *tw[2] = {&f->d_twA, &f->d_twB};
    for (int p = 0; p < 2; p++) {
        std::vector<float2> t;
        cv[p]->fill_tw(t);
        CLB_TRY(tw[p]->reserve(t.size() * sizeof(float2)));
        CLB_CUDA(cudaMemcpy(tw[p]->p, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
        const void *k = (const void *)cv[p]->kernel[p];
        int occ = 0;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cv[p]->smem_bytes) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, cv[p]->threads, cv[p]->smem_bytes) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            return CLB200_OK;                  // does not fit: the caller falls back to the four-step path
        }
        f->col_resident[p] = occ;
    }
    {
        const int nhi = std::max(1, f->n / 512);
        std::vector<float2> t(512 + nhi, make_float2(1.f, 0.f));
        for (int j = 0; j < 512; j++) {
            const double a = -2.0 * M_PI * (double)j / (double)f->n;
            t[j] = make_float2((float)cos(a), (float)sin(a));
        }
        for (int j = 0; j < nhi; j++) {
            const double b = -2.0 * M_PI * (512.0 * j) / (double)f->n;
            t[512 + j] = make_float2((float)cos(b), (float)sin(b));
        }
        CLB_TRY(f->d_tw4.reserve(t.size() * sizeof(float2)));
        CLB_CUDA(cudaMemcpy(f->d_tw4.p, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    f->colA = ca;
    f->colB = cb;
    f->two_pass = true;
    return CLB200_OK;
}

} // namespace

extern "C" {

int clb200_xcorr_fft_create(int fft_size, int num_inputs, int input_type, int device, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(fft_size >= 2 && (fft_size & (fft_size - 1)) == 0 && fft_size <= 16384, CLB200_EINVAL,
              "clxcorrelate_fft_vcf: fft size %d is not a power of two in 2..16384", fft_size);
    CLB_CHECK(num_inputs >= 2 && num_inputs <= CLB200_XCFFT_MAX_INPUTS, CLB200_EINVAL,
              "clxcorrelate_fft_vcf: 2..%d inputs, got %d", CLB200_XCFFT_MAX_INPUTS, num_inputs);
    CLB_CHECK(input_type == 1 || input_type == 2, CLB200_EINVAL,
              "clxcorrelate_fft_vcf: input type must be 1 (FFT) or 2 (time series), got %d", input_type);
    int nd = clb200_device_count();
    CLB_CHECK(nd > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < nd, CLB200_EINVAL, "device %d out of range", device);
    DeviceGuard g(device);
    XcFft *x = new XcFft;
    x->kind = KIND_XCFFT;
    x->device = device;
    x->n = fft_size;
    x->num_inputs = num_inputs;
    x->input_type = input_type;
    x->var = pick_variant(ilog2(fft_size));
    auto fail = [&](int rc) {
        delete x;
        return rc;
    };
    std::vector<float2> tw;
    x->var->fill_tw(tw);
    if (x->d_tw.reserve(tw.size() * sizeof(float2)) != CLB200_OK) return fail(CLB200_ENOMEM);
    if (cudaMemcpy(x->d_tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("clxcorrelate_fft_vcf: twiddle upload failed");
        return fail(CLB200_ECUDA);
    }
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute((const void *)x->var->kernel_xc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         x->var->smem_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void *)x->var->kernel[0], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 x->var->smem_bytes);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)x->var->kernel_xc, x->var->threads,
                                                          x->var->smem_bytes);
    if (e != cudaSuccess || occ < 1) {
        set_error("clxcorrelate_fft_vcf: kernel for size %d does not fit an SM (%s)", fft_size, cudaGetErrorString(e));
        return fail(CLB200_ECUDA);
    }
    x->resident = occ;
    *out = x;
    return CLB200_OK;
}

int clb200_xcorr_fft_launch_device(clb200_handle h, const void *const *d_in, void *const *d_out, long nvec,
                                   void *stream)
{
    XcFft *x;
    CLB_TRY(check_kind(h, KIND_XCFFT, &x));
    CLB_CHECK(d_in && d_out && nvec >= 0, CLB200_EINVAL, "bad arguments");
    DeviceGuard g(x->device);
    return xcfft_launch(x, d_in, d_out, nvec, (cudaStream_t)stream);
}

int clb200_xcorr_fft_work(clb200_handle h, const void *const *in, void *const *out, long nvec)
{
    XcFft *x;
    CLB_TRY(check_kind(h, KIND_XCFFT, &x));
    CLB_CHECK(in && out && nvec >= 0, CLB200_EINVAL, "bad arguments");
    if (nvec == 0) return CLB200_OK;
    DeviceGuard g(x->device);
    if (!x->st) CLB_CUDA(cudaStreamCreateWithFlags(&x->st, cudaStreamNonBlocking));
    const size_t ib = (size_t)nvec * x->n * sizeof(float2), ob = (size_t)nvec * x->n * sizeof(float);
    CLB_TRY(x->d_in.reserve(ib * x->num_inputs));
    CLB_TRY(x->d_out.reserve(ob * (x->num_inputs - 1)));
    const void *di[CLB200_XCFFT_MAX_INPUTS];
    void *dout[CLB200_XCFFT_MAX_INPUTS];
    for (int k = 0; k < x->num_inputs; k++) {
        CLB_CHECK(in[k] != nullptr, CLB200_EINVAL, "null input %d", k);
        di[k] = (char *)x->d_in.p + ib * k;
        CLB_CUDA(cudaMemcpyAsync((void *)di[k], in[k], ib, cudaMemcpyHostToDevice, x->st));
        x->n_h2d += ib;
    }
    for (int k = 0; k + 1 < x->num_inputs; k++) dout[k] = (char *)x->d_out.p + ob * k;
    CLB_TRY(xcfft_launch(x, di, dout, nvec, x->st));
    for (int k = 0; k + 1 < x->num_inputs; k++) {
        CLB_CHECK(out[k] != nullptr, CLB200_EINVAL, "null output %d", k);
        CLB_CUDA(cudaMemcpyAsync(out[k], dout[k], ob, cudaMemcpyDeviceToHost, x->st));
        x->n_d2h += ob;
    }
    CLB_CUDA(cudaStreamSynchronize(x->st));
    return CLB200_OK;
}


int clb200_fft_create(int fft_size, int dir, const float *window, int window_len, int dtype,
                      int device, int shift, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    const bool pow2 = fft_size >= 2 && (fft_size & (fft_size - 1)) == 0;
    CLB_CHECK(fft_size >= 2 && fft_size <= (pow2 ? (1 << 22) : (1 << 21)), CLB200_EINVAL,
              "clFFT: fft size %d is outside 2..4194304 (powers of two) / 2..2097152 (other lengths)", fft_size);
    CLB_CHECK(dir == CLB200_FFT_FORWARD || dir == CLB200_FFT_BACKWARD, CLB200_EINVAL,
              "clFFT: direction must be -1 (forward) or 1 (backward), got %d", dir);
    // lib/clFFT_impl.cc:74-76: "window not the same length as fft_size"
    CLB_CHECK(window_len == 0 || window_len == fft_size, CLB200_EINVAL,
              "clFFT: window not the same length as fft_size (%d vs %d)", window_len, fft_size);
    CLB_CHECK(window_len == 0 || window != nullptr, CLB200_EINVAL, "clFFT: null window");
    CLB_CHECK(dtype == CLB200_DTYPE_COMPLEX || dtype == CLB200_DTYPE_FLOAT, CLB200_EINVAL,
              "clFFT: data type must be complex or float, got %d", dtype);
    CLB_CHECK(dtype == CLB200_DTYPE_COMPLEX || dir == CLB200_FFT_FORWARD, CLB200_EINVAL,
              "clFFT: real input is forward-only");
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);

    DeviceGuard g(device);
    Fft *f = new Fft;
    f->kind = KIND_FFT;
    f->device = device;
    f->init_work_counters();
    f->n = fft_size;
    f->logn = ilog2(fft_size);
    f->dir = dir;
    f->dtype = dtype;
    // the reference shifts complex data only: a real-input spectrum is left unshifted (lib/clFFT_impl.cc:594)
    f->shift = (shift && dtype == CLB200_DTYPE_COMPLEX) ? 1 : 0;
    f->mode = dtype == CLB200_DTYPE_FLOAT ? 2 : (dir > 0 ? 1 : 0);
    auto fail = [&](int rc) {
        delete f;
        return rc;
    };
    if (!pow2) {
        // Bluestein: m = power of two >= 2n - 1
        f->blue = true;
        f->logn = 0;
        int m = 1;
        while (m < 2 * fft_size - 1) m <<= 1;
        f->m = m;
        if (window_len) {
            f->has_window = true;
            if (f->d_win.reserve(sizeof(float) * fft_size) != CLB200_OK) return fail(CLB200_ENOMEM);
            if (cudaMemcpy(f->d_win.p, window, sizeof(float) * fft_size, cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("clFFT: window upload failed");
                return fail(CLB200_ECUDA);
            }
        }
        // chirp in double with the phase reduced exactly: n^2 mod 2N
        std::vector<float2> w(fft_size);
        std::vector<double> bre(m, 0.0), bim(m, 0.0);
        for (int i = 0; i < fft_size; i++) {
            const long q = ((long)i * i) % (2L * fft_size);
            const double a = -M_PI * (double)q / (double)fft_size;
            w[i] = make_float2((float)cos(a), (float)sin(a));
            bre[i] = cos(a);
            bim[i] = -sin(a);                    // conj(w)
            if (i) {
                bre[m - i] = bre[i];
                bim[m - i] = bim[i];
            }
        }
        // spectrum of b by a double-precision radix-2 FFT on the host (once per plan), scaled by 1/m
        {
            int lg = 0;
            while ((1 << lg) < m) lg++;
            for (int i = 0; i < m; i++) {
                int r = 0;
                for (int b = 0; b < lg; b++)
                    if (i & (1 << b)) r |= 1 << (lg - 1 - b);
                if (r > i) {
                    std::swap(bre[i], bre[r]);
                    std::swap(bim[i], bim[r]);
                }
            }
            for (int len = 2; len <= m; len <<= 1) {
                const int half = len >> 1;
                for (int k = 0; k < half; k++) {
                    const double a = -2.0 * M_PI * (double)k / (double)len, wr = cos(a), wi = sin(a);
                    for (int i = k; i < m; i += len) {
                        const int j = i + half;
                        const double tr = bre[j] * wr - bim[j] * wi, ti = bre[j] * wi + bim[j] * wr;
                        bre[j] = bre[i] - tr;
                        bim[j] = bim[i] - ti;
                        bre[i] += tr;
                        bim[i] += ti;
                    }
                }
            }
        }
        std::vector<float2> bs(m);
        for (int i = 0; i < m; i++) bs[i] = make_float2((float)(bre[i] / m), (float)(bim[i] / m));
        if (f->d_chirp.reserve(sizeof(float2) * fft_size) != CLB200_OK || f->d_bspec.reserve(sizeof(float2) * m) != CLB200_OK)
            return fail(CLB200_ENOMEM);
        if (cudaMemcpy(f->d_chirp.p, w.data(), sizeof(float2) * fft_size, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(f->d_bspec.p, bs.data(), sizeof(float2) * m, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("clFFT: chirp upload failed");
            return fail(CLB200_ECUDA);
        }
        int rc = clb200_fft_create(m, CLB200_FFT_FORWARD, nullptr, 0, CLB200_DTYPE_COMPLEX, device, 0, &f->sub1);
        if (rc == CLB200_OK) rc = clb200_fft_create(m, CLB200_FFT_BACKWARD, nullptr, 0, CLB200_DTYPE_COMPLEX, device, 0, &f->sub2);
        if (rc != CLB200_OK) return fail(rc);
        {
            // both stages fused into the two in-SM transforms when m fits one CTA (CLB200_FFT_CZ_UNFUSED=1 for A/B)
            const char *e = getenv("CLB200_FFT_CZ_UNFUSED");
            Fft *s1 = static_cast<Fft *>(f->sub1);
            if (!(e && atoi(e)) && s1->var) {
                int occ0 = 0, occ1 = 0;
                const void *k0 = (const void *)s1->var->kernel_cz[0], *k1 = (const void *)s1->var->kernel_cz[1];
                if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, s1->var->smem_bytes) == cudaSuccess &&
                    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, s1->var->smem_bytes) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, k0, s1->var->threads, s1->var->smem_bytes) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, s1->var->threads, s1->var->smem_bytes) == cudaSuccess &&
                    occ0 >= 1 && occ1 >= 1) {
                    f->blue_fused = true;
                    f->blue_resident = std::min(occ0, occ1);
                }
                cudaGetLastError();
            }
        }
        f->set_info("clFFT %d-pt %s: not a power of two -> chirp-z (Bluestein) over %d-point plans, %s", fft_size,
                    dir < 0 ? "forward" : "backward", m,
                    f->blue_fused ? "pre-multiply / zero-pad fused into the forward transform, spectrum product and post-multiply "
                                    "into the inverse one (two kernels)"
                                  : "window / half swaps fused into the pre- and post-multiply kernels (five kernels)");
        *out = f;
        return CLB200_OK;
    }
    if (fft_size > 16384) {
        // four-step path: two plans of the in-SM kernels (n1 >= n2, both <= 16384 up to 2^28) + transposes
        f->big = true;
        f->n1 = 1 << ((f->logn + 1) / 2);
        f->n2 = fft_size / f->n1;
        if (window_len) {
            f->has_window = true;
            if (f->d_win.reserve(sizeof(float) * fft_size) != CLB200_OK) return fail(CLB200_ENOMEM);
            if (cudaMemcpy(f->d_win.p, window, sizeof(float) * fft_size, cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("clFFT: window upload failed");
                return fail(CLB200_ECUDA);
            }
        }
        {
            const char *e = getenv("CLB200_FFT_FOURSTEP");          // A/B: force the five-pass path
            const ColVariant *ca = pick_col(ilog2(f->n1)), *cb = pick_col(ilog2(f->n2));
            if (!(e && atoi(e)) && ca && cb) {
                int rc2 = fft_setup_two_pass(f, ca, cb);
                if (rc2 != CLB200_OK) return fail(rc2);
                if (f->two_pass) {
                    f->set_info("clFFT %d-pt %s: two passes over HBM: %d-point column transforms (window / swaps / twiddles fused, "
                                "stored along k1) and %d-point column transforms, 16 columns per CTA", fft_size,
                                dir < 0 ? "forward" : "backward", f->n1, f->n2);
                    *out = f;
                    return CLB200_OK;
                }
            }
        }
        int rc = clb200_fft_create(f->n1, dir, nullptr, 0, CLB200_DTYPE_COMPLEX, device, 0, &f->sub1);
        if (rc == CLB200_OK) rc = clb200_fft_create(f->n2, dir, nullptr, 0, CLB200_DTYPE_COMPLEX, device, 0, &f->sub2);
        if (rc != CLB200_OK) return fail(rc);
        f->set_info("clFFT %d-pt %s: four-step, %d x %d-pt and %d x %d-pt in-SM passes around three tiled transposes "
                    "(window / half swaps / twiddles fused into the transposes)", fft_size, dir < 0 ? "forward" : "backward",
                    f->n2, f->n1, f->n1, f->n2);
        *out = f;
        return CLB200_OK;
    }
    f->var = pick_variant(f->logn);
    if (!f->var) {
        set_error("clFFT: no kernel for size %d", fft_size);
        return fail(CLB200_EINVAL);
    }
    std::vector<float2> tw;
    f->var->fill_tw(tw);
    if (f->d_tw.reserve(tw.size() * sizeof(float2)) != CLB200_OK) return fail(CLB200_ENOMEM);
    if (cudaMemcpy(f->d_tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice) !=
        cudaSuccess) {
        set_error("clFFT: twiddle upload failed");
        return fail(CLB200_ECUDA);
    }
    if (window_len) {
        f->has_window = true;
        if (f->d_win.reserve(sizeof(float) * fft_size) != CLB200_OK) return fail(CLB200_ENOMEM);
        if (cudaMemcpy(f->d_win.p, window, sizeof(float) * fft_size, cudaMemcpyHostToDevice) !=
            cudaSuccess) {
            set_error("clFFT: window upload failed");
            return fail(CLB200_ECUDA);
        }
    }
    cudaError_t e = cudaFuncSetAttribute((const void *)f->var->kernel[f->mode],
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         f->var->smem_bytes);
    if (e != cudaSuccess) {
        set_error("clFFT: cannot reserve %d B of shared memory: %s", f->var->smem_bytes,
                  cudaGetErrorString(e));
        return fail(CLB200_ECUDA);
    }
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)f->var->kernel[f->mode],
                                                      f->var->threads, f->var->smem_bytes);
    if (e != cudaSuccess || occ < 1) {
        set_error("clFFT: kernel for size %d does not fit an SM (%s)", fft_size,
                  cudaGetErrorString(e));
        return fail(CLB200_ECUDA);
    }
    f->resident = occ;
    f->loop_form = fft_loop_form(f->logn);
    if (f->loop_form) {
        const void *k = (const void *)(f->loop_form == 2 ? f->var->kernel_u[f->mode] : f->var->kernel_d[f->mode]);
        int occ2 = 0;
        if (k && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, f->var->smem_bytes) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k, f->var->threads, f->var->smem_bytes) == cudaSuccess &&
            occ2 >= occ)
            f->dyn_ok = true;
        cudaGetLastError();
    }
    {
        // pass-1 twiddle table in shared memory (complex modes, tables up to 8 KiB) when it keeps the occupancy
        const char *ts = getenv("CLB200_FFT_TW1S");
        const int want = ts ? atoi(ts) : 0;
        if (want && f->mode < 2 && f->var->tw1_bytes > 0 && f->var->tw1_bytes <= 8192) {
            const void *k = (const void *)f->var->kernel_s[f->mode];
            int occ2 = 0;
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, f->var->smem_bytes + f->var->tw1_bytes) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k, f->var->threads, f->var->smem_bytes + f->var->tw1_bytes) == cudaSuccess &&
                occ2 >= occ)
                f->use_tw1s = true;
            cudaGetLastError();
        }
    }
    // opt-in: measured on B200 it is on par (32 elements/thread) or slower (16) than the
    // plain kernel at 8192 points -- two resident CTAs already overlap each other's loads
    const char *pfenv = getenv("CLB200_FFT_PREFETCH");
    if (f->mode < 2 && f->var->kernel_pf[f->mode] && pfenv && atoi(pfenv)) {
        const void *k = (const void *)f->var->kernel_pf[f->mode];
        int occ2 = 0;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, f->var->smem_bytes + 16) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k, f->var->threads, f->var->smem_bytes + 16) == cudaSuccess &&
            occ2 >= occ)
            f->use_pf = true;
        cudaGetLastError();
    }
    f->set_info("clFFT %d-pt %s%s%s%s: k_fft<%d,%d,%d,...> %d threads, %d transform(s)/CTA, %d B shared memory, %d CTAs/SM -> persistent grid of %d",
                fft_size, dir < 0 ? "forward" : "backward", dtype == CLB200_DTYPE_FLOAT ? ", real input" : "",
                f->has_window ? ", window fused" : "", f->shift ? ", half swap fused" : "", f->var->logn, f->var->ept, f->var->batch,
                f->var->threads, f->var->batch, f->var->smem_bytes, f->resident, f->resident * device_sm_count(device));
    *out = f;
    return CLB200_OK;
}

int clb200_fft_launch_device(clb200_handle h, const void *d_in, void *d_out, long nvec, void *stream)
{
    Fft *f;
    CLB_TRY(check_kind(h, KIND_FFT, &f));
    CLB_CHECK(nvec >= 0, CLB200_EINVAL, "negative vector count");
    DeviceGuard g(f->device);
    return fft_launch(f, d_in, d_out, nvec, (cudaStream_t)stream);
}

int clb200_fft_work(clb200_handle h, const void *in, void *out, long nvec)
{
    Fft *f;
    CLB_TRY(check_kind(h, KIND_FFT, &f));
    CLB_CHECK(nvec >= 0, CLB200_EINVAL, "negative vector count");
    if (nvec == 0) return CLB200_OK;
    DeviceGuard g(f->device);
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = (size_t)f->n * (f->dtype == CLB200_DTYPE_FLOAT ? 4 : 8);
    pd.out_bytes[0] = (size_t)f->n * 8;
    return run_chunked(f, pd, nvec, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           return fft_launch(f, di[0], dout[0], n, st);
                       });
}

int clb200_fft_work_streams(clb200_handle h, const void *const *in, void *const *out, int nstreams,
                            long nvec)
{
    CLB_CHECK(nstreams >= 1 && in && out, CLB200_EINVAL, "bad stream list");
    for (int s = 0; s < nstreams; s++) CLB_TRY(clb200_fft_work(h, in[s], out[s], nvec));
    return CLB200_OK;
}

} // extern "C"
