// fft_device.cuh -- device building blocks of the hand-written FFT.
//
// The arithmetic that clFFT (clfftEnqueueTransform, lib/clFFT_impl.cc:583) and
// FFTW (lib/fft.cc:175-179) do for the reference is done here by a Stockham
// autosort FFT:  N = R0*R1*...; pass p with Ns = R0*..*R(p-1):
//     for j in [0, N/R):  k = j mod Ns
//         v[r] = x[j + r*N/R] * exp(-2*pi*i * r*k / (Ns*R))        r = 0..R-1
//         V    = DFT_R(v)
//         y[(j-k)*R + k + r*Ns] = V[r]
// Radix-R butterflies (R = 2..32) run entirely in registers as unrolled
// radix-2 DIF stages with compile-time twiddles; passes exchange data through
// padded shared memory; the first pass reads HBM directly (coalesced float2)
// and the last pass writes HBM directly, so each sample crosses HBM exactly
// once in each direction (16 B/sample).
//
// The inverse transform is the forward one on data with re/im swapped on load
// and on store (IDFT(x) = swap(DFT(swap(x)))), so only forward twiddles exist.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <utility>

namespace clb200 {
namespace fftdev {

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

__host__ __device__ constexpr int bitrev(int v, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; b++)
        if (v & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}

// cos(2*pi*k/32), k = 0..8, correctly rounded literals
__host__ __device__ constexpr float cos32_q(int k)
{
    constexpr float c[9] = {1.0f,
                            0.98078528040323044913f,
                            0.92387953251128675613f,
                            0.83146961230254523708f,
                            0.70710678118654752440f,
                            0.55557023301960222474f,
                            0.38268343236508977173f,
                            0.19509032201612826785f,
                            0.0f};
    return c[k];
}
__host__ __device__ constexpr float cos32(int k)
{
    k = ((k % 32) + 32) % 32;
    if (k > 16) k = 32 - k;
    return k > 8 ? -cos32_q(16 - k) : cos32_q(k);
}
__host__ __device__ constexpr float sin32(int k) { return cos32(k - 8); }

// a * W_32^K  (W = exp(-2*pi*i/32)), K compile-time
template <int K>
__device__ __forceinline__ float2 mul_w32(float2 a)
{
    constexpr int k = ((K % 32) + 32) % 32;
    if constexpr (k == 0) return a;
    else if constexpr (k == 8) return make_float2(a.y, -a.x);
    else if constexpr (k == 16) return make_float2(-a.x, -a.y);
    else if constexpr (k == 24) return make_float2(-a.y, a.x);
    else if constexpr (k == 4) {
        constexpr float s = cos32(4);
        return make_float2((a.x + a.y) * s, (a.y - a.x) * s);
    } else if constexpr (k == 12) {
        constexpr float s = cos32(4);
        return make_float2((a.y - a.x) * s, -(a.x + a.y) * s);
    } else {
        constexpr float wr = cos32(k), wi = -sin32(k);
        return make_float2(fmaf(-a.y, wi, a.x * wr), fmaf(a.y, wr, a.x * wi));
    }
}

// Complex add / subtract as ONE packed instruction (add.f32x2 -> FADD2 on sm_100a) instead of two FADDs.
// The FP32 pipe does the same lane-cycles either way (measured: 64 complex adds / clock / SM for both),
// but a butterfly-heavy kernel issues half as many instructions for its adds -- which are half of
// all instructions of an FFT -- and the freed issue slots go to the loads, stores and exchanges.
#ifndef CLB_F32X2
#define CLB_F32X2 1
#endif
// Twiddle loads traded for arithmetic (the FFT kernels are co-limited by L1 wavefronts and load-queue
// stalls while the FP32 pipe is under 40 % busy):
//   1: last pass, several sub-blocks per thread -- one table load per r, the other sub-blocks rotate it by a
//      compile-time W_32 power (8192 points: 59 -> 38 twiddle loads per thread, 4.92 -> 5.17 TB/s);
//   2: additionally only r = 1, 2, 4, ... are loaded and the other W^(r*k) are products of those (one to four
//      extra roundings, error unchanged at 2e-7): 4096 points 5.14 -> 5.76 TB/s, 8192 +0.4 %.  Used by fft.cu;
//      the one-warp FFT filter is slower with it (3.66 -> 3.50 TB/s) and stays at 1.
#ifndef CLB_TW_DERIVE
#define CLB_TW_DERIVE 1
#endif
__device__ __forceinline__ float2 cadd(float2 a, float2 c)
{
#if CLB_F32X2
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
#else
    return make_float2(a.x + c.x, a.y + c.y);
#endif
}
__device__ __forceinline__ float2 csub(float2 a, float2 c)
{
#if CLB_F32X2
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
#else
    return make_float2(a.x - c.x, a.y - c.y);
#endif
}

// acc + (t, t) * v as one packed fma (FFMA2): tt must hold the same real factor in both halves
__device__ __forceinline__ float2 cfma_real(float2 tt, float2 v, float2 acc)
{
#if CLB_F32X2
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&tt)),
        "l"(*reinterpret_cast<unsigned long long *>(&v)), "l"(*reinterpret_cast<unsigned long long *>(&acc)));
    return *reinterpret_cast<float2 *>(&r);
#else
    return make_float2(fmaf(tt.x, v.x, acc.x), fmaf(tt.y, v.y, acc.y));
#endif
}

// In-register forward DFT of x[OFF .. OFF+R): radix-2 decimation in frequency.
// Result is left in bit-reversed order: X[k] = x[OFF + bitrev(k)].
template <int R, int OFF, int NREG>
__device__ __forceinline__ void dft_dif(float2 (&x)[NREG])
{
    constexpr int LR = ilog2(R);
    static_for<0, LR>([&](auto s_) {
        constexpr int s = decltype(s_)::value;
        constexpr int h = R >> (s + 1);
        static_for<0, R / 2>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            constexpr int g = b / h, i = b % h;
            constexpr int ia = OFF + g * 2 * h + i, ib = ia + h;
            float2 a = x[ia], c = x[ib];
            x[ia] = cadd(a, c);
            float2 d = csub(a, c);
            x[ib] = mul_w32<i * (16 / h)>(d);
        });
    });
}

// In-register forward DFT, radix-2 decimation in TIME: input sample n sits in x[OFF + bitrev(n)], the result is
// left in natural order (X[k] = x[OFF + k]).  Same operation count and instruction mix as dft_dif (the rotation is
// applied to the second operand before the add instead of to the difference after it); a DIF pass followed by a DIT
// pass takes natural-order registers to natural-order registers, so two transforms can run through ONE copy of the
// code in a loop (the FFT filter's instruction footprint, filter.cu).
template <int R, int OFF, int NREG>
__device__ __forceinline__ void dft_dit(float2 (&x)[NREG])
{
    constexpr int LR = ilog2(R);
    static_for<0, LR>([&](auto s_) {
        constexpr int s = decltype(s_)::value;
        constexpr int h = 1 << s;
        static_for<0, R / 2>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            constexpr int g = b / h, i = b % h;
            constexpr int ia = OFF + g * 2 * h + i, ib = ia + h;
            float2 a = x[ia], c = mul_w32<i * (16 / h)>(x[ib]);
            x[ia] = cadd(a, c);
            x[ib] = csub(a, c);
        });
    });
}

// streaming 64-bit load that does not allocate in L1 (L1 is kept for the twiddle tables)
__device__ __forceinline__ float2 ldg_stream(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy, bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.y, w.x, a.x * w.y));
}

// shared-memory padding: 16 B after every 128 B (float2 units) keeps the
// stride-R stores of the first pass and the unit-stride loads conflict-free
// (granularity G = 16 float2; 32 when the first radix is 32, whose 256 B per-thread
// runs need an odd multiple of 16 B as thread stride)
template <int G>
__host__ __device__ constexpr int padg(int i) { return i + ((i / G) << 1); }

// Compile-time radix plan: greedy, largest radices first.
template <int LOGN, int EPT>
struct Plan {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = N / EPT;             // threads per transform
    __host__ __device__ static constexpr int npass()
    {
        int rem = N, n = 0;
        while (rem > 1) {
            rem /= (rem < EPT ? rem : EPT);
            n++;
        }
        return n;
    }
    __host__ __device__ static constexpr int ns(int p)      // product of earlier radices
    {
        int rem = N, s = 1;
        for (int i = 0; i < p; i++) {
            int r = rem < EPT ? rem : EPT;
            rem /= r;
            s *= r;
        }
        return s;
    }
    __host__ __device__ static constexpr int radix(int p)
    {
        int rem = N / ns(p);
        return rem < EPT ? rem : EPT;
    }
    __host__ __device__ static constexpr int tw_offset(int p)   // float2 units, passes >= 1
    {
        int off = 0;
        for (int i = 1; i < p; i++) off += (radix(i) - 1) * ns(i);
        return off;
    }
    static constexpr int TW_TOTAL = tw_offset(npass());
    static constexpr int PADG = EPT >= 32 ? 32 : 16;             // padding granularity (float2)
    __host__ __device__ static constexpr int pad(int i) { return padg<PADG>(i); }
    static constexpr int SMEM_F2 = npass() > 1 ? pad(N) : 0;    // float2 per transform
};

// ---------------------------------------------------------------------------
// fft_core: all passes of one transform for one thread.
//   in : x[u*R0 + r]  = element  lt + u*T + r*(N/R0)       (first-pass order)
//   out: x[u*RL + bitrev(r)] = element (j-k)*RL + k + r*NSL, j = lt+u*T, k = j mod NSL
//        (for the greedy plan NSL = N/RL, so that is  j + r*(N/RL) )
// buf is this transform's padded shared-memory line (unused when npass == 1).
// Barriers are CTA-wide: every thread of the CTA must call this the same
// number of times.  The first barrier protects buf against readers of the
// previous use (previous transform's last pass).
// ---------------------------------------------------------------------------
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// after_last_load() runs once, right after the last pass has read its inputs out of
// shared memory (the buffer is dead from then on for this transform) -- the
// prefetching kernel uses it to start the bulk copy of the next vector.
// tw1s (optional): a shared-memory copy of the pass-1 twiddle table (P::tw_offset(1) .. tw_offset(2)); its
// entries are the same for every transform a thread runs, and reading them with LDS keeps them out of the
// global-load queue (the 8192-point kernel stalls on lg_throttle: 59 twiddle LDGs per 32 data LDGs).
// TWD: twiddle derivation mode (see CLB_TW_DERIVE)
// V128 = false: no 128-bit shared-memory stores (lines that are only 8 B aligned: the column transforms of fft.cu)
template <class P, int EPT, class Hook = NoHook, int TWD = CLB_TW_DERIVE, bool V128 = true>
__device__ __forceinline__ void fft_core(float2 (&x)[EPT], float2 *buf, int lt,
                                         const float2 *__restrict__ tw, Hook after_last_load = Hook{},
                                         const float2 *tw1s = nullptr)
{
    constexpr int N = P::N, T = P::T, NPASS = P::npass();
    // P::pad(a + c) == P::pad(a) + P::pad(c) whenever c is a multiple of the granularity, so nearly every
    // shared-memory address below is one per-thread base plus a compile-time offset
    float2 *const ldp = buf + P::pad(lt);
    static_for<0, NPASS>([&](auto p_) {
        constexpr int p = decltype(p_)::value;
        constexpr int R = P::radix(p), NS = P::ns(p), NB = EPT / R;
        constexpr bool first = (p == 0), last = (p == NPASS - 1);
        constexpr int STR = N / R;
        constexpr int LR = ilog2(R);

        if constexpr (!first) {
            static_for<0, NB * R>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                constexpr int c = (e / R) * T + (e % R) * STR;
                if constexpr (c % P::PADG == 0) x[e] = ldp[P::pad(c)];
                else x[e] = buf[P::pad(lt + c)];
            });
            if constexpr (!last) __syncthreads();     // all reads before the in-place writes
            else after_last_load();
        }

        // last pass with several sub-blocks per thread: k = lt + u*T, so W^(r*k) = W^(r*lt) * W_32^(r*u*32/EPT):
        // one table load per r, the sub-blocks u > 0 rotate it by a compile-time constant
        constexpr bool derive = TWD >= 1 && !first && last && NB > 1 && R >= 4 && NS * R == N;   // (radix 2: three loads saved do not pay)
        if constexpr (derive) {
            const float2 *t = tw + P::tw_offset(p) + lt;
            float2 wp[R] = {};
            static_for<1, R>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                if constexpr (TWD >= 2 && N >= 512 && (r & (r - 1)) != 0) wp[r] = cmul(wp[r & (r - 1)], wp[r & -r]);
                else wp[r] = __ldg(t + (r - 1) * NS);
                static_for<0, NB>([&](auto u_) {
                    constexpr int u = decltype(u_)::value;
                    x[u * R + r] = cmul(x[u * R + r], mul_w32<(r * u * (32 / EPT)) % 32>(wp[r]));
                });
            });
        }
        static_for<0, NB>([&](auto u_) {
            constexpr int u = decltype(u_)::value;
            if constexpr (!first && !derive) {
                const int k = (lt + u * T) & (NS - 1);
                if (p == 1 && tw1s != nullptr) {
                    const float2 *t = tw1s + k;
#pragma unroll
                    for (int r = 1; r < R; r++) x[u * R + r] = cmul(x[u * R + r], t[(r - 1) * NS]);
                } else if constexpr (TWD >= 2 && R >= 8 && N >= 512) {       // (64 / 256 points: -1.5 % with it)
                    // W^(r*k) for r = 1, 2, 4, ... from the table, the other r as products of those
                    const float2 *t = tw + P::tw_offset(p) + k;
                    float2 wp[R] = {};
                    static_for<1, R>([&](auto r_) {
                        constexpr int r = decltype(r_)::value;
                        if constexpr ((r & (r - 1)) != 0) wp[r] = cmul(wp[r & (r - 1)], wp[r & -r]);
                        else wp[r] = __ldg(t + (r - 1) * NS);
                        x[u * R + r] = cmul(x[u * R + r], wp[r]);
                    });
                } else {
                    const float2 *t = tw + P::tw_offset(p) + k;
#pragma unroll
                    for (int r = 1; r < R; r++) x[u * R + r] = cmul(x[u * R + r], __ldg(t + (r - 1) * NS));
                }
            }
            dft_dif<R, u * R>(x);
        });

        if constexpr (!last) {
            if constexpr (first) __syncthreads();
            if constexpr (V128 && NS == 1 && (R % 2 == 0)) {
                // thread-contiguous run of R outputs: 128-bit stores
                float2 *const sp = buf + P::pad(lt * R);
                static_for<0, NB>([&](auto u_) {
                    constexpr int u = decltype(u_)::value;
                    static_for<0, R / 2>([&](auto q_) {
                        constexpr int q = decltype(q_)::value;
                        float2 a = x[u * R + bitrev(2 * q, LR)], b = x[u * R + bitrev(2 * q + 1, LR)];
                        float2 *d;
                        if constexpr ((u * T * R) % P::PADG == 0) d = sp + P::pad(u * T * R + 2 * q);
                        else d = buf + P::pad((lt + u * T) * R + 2 * q);
                        *reinterpret_cast<float4 *>(d) = make_float4(a.x, a.y, b.x, b.y);
                    });
                });
            } else {
                const int k0 = lt & (NS - 1);
                float2 *const sp = buf + P::pad((lt - k0) * R + k0);
                static_for<0, NB>([&](auto u_) {
                    constexpr int u = decltype(u_)::value;
                    static_for<0, R>([&](auto r_) {
                        constexpr int r = decltype(r_)::value;
                        // j = lt + u*T keeps k when NS divides T; the tile moves by u*T*R
                        if constexpr ((u * T) % NS == 0 && (u * T * R) % P::PADG == 0 && NS % P::PADG == 0) {
                            sp[P::pad(u * T * R + r * NS)] = x[u * R + bitrev(r, LR)];
                        } else {
                            const int j = lt + u * T, k = j & (NS - 1);
                            buf[P::pad((j - k) * R + k + r * NS)] = x[u * R + bitrev(r, LR)];
                        }
                    });
                });
            }
            __syncthreads();
        }
    });
}

// element index of first-pass register slot e (= u*R0 + r)
template <class P, int EPT>
__device__ __forceinline__ constexpr int in_index(int lt, int e)
{
    constexpr int R0 = P::radix(0);
    return lt + (e / R0) * P::T + (e % R0) * (P::N / R0);
}

// f(natural_index, value) for every output this thread holds after fft_core
template <class P, int EPT, class F>
__device__ __forceinline__ void for_each_output(const float2 (&x)[EPT], int lt, F &&f)
{
    constexpr int NPASS = P::npass();
    constexpr int R = P::radix(NPASS - 1), NS = P::ns(NPASS - 1), NB = EPT / R, LR = ilog2(R);
    static_for<0, NB>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        const int j = lt + u * P::T;
        const int k = j & (NS - 1);
        const int base = (j - k) * R + k;
        static_for<0, R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            f(base + r * NS, x[u * R + bitrev(r, LR)]);
        });
    });
}

// same, but the index is handed over as (lt + compile-time offset): for the greedy
// plan the last pass has NS = N/R, so output (u, r) is element lt + u*T + r*NS
template <class P, int EPT, class F>
__device__ __forceinline__ void for_each_output_c(const float2 (&x)[EPT], F &&f)
{
    constexpr int NPASS = P::npass();
    constexpr int R = P::radix(NPASS - 1), NS = P::ns(NPASS - 1), NB = EPT / R, LR = ilog2(R);
    static_assert(NS * R == P::N, "last pass of the greedy plan spans the transform");
    static_for<0, NB>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        static_for<0, R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            f(std::integral_constant<int, u * P::T + r * NS>{}, x[u * R + bitrev(r, LR)]);
        });
    });
}

} // namespace fftdev
} // namespace clb200
