// filter.cu -- clFilter: complex data, real taps; time-domain FIR and FFT filter.
//
// Reference:
//   time domain  clFilter_impl::filterGPUTimeDomain + td_FIR_complex
//                (lib/clFilter_impl.cc:505-589, :162-194):
//                out[g] = sum_{i<K} taps[K-1-i] * in[g+i], history K-1 in front.
//   freq domain  clFilter_impl::filterGPUFrequencyDomain (:592-681) on top of
//                fft_filter_ccf (lib/fft_filter.cc:38-97,133-175): per block of
//                nsamples: H2D, clFFT forward, D2H, CPU multiply by the tap
//                spectrum, H2D, clFFT inverse, D2H, CPU tail add -- four PCIe
//                crossings per 257 samples at 256 taps.
// Here the frequency-domain path is ONE kernel: a CTA loads NF = L + K - 1
// stream samples (overlap-save: the K-1 overlap is re-read from L2, nothing is
// zero padded), runs the forward FFT in registers/shared memory, multiplies by
// the resident tap spectrum (pre-scaled by 1/NF like fft_filter.cc:52), runs
// the inverse FFT and stores the L valid outputs (decimated) -- the samples cross
// HBM once each way (16 B/sample algorithmic).  When the first and last radix
// of the plan are equal the spectrum never leaves registers between the two
// transforms.
//
// Stream state (the reference's set_history(K) + d_tail + dec_ctr): the last K-1
// input samples and the decimation phase live in the handle, device-resident.
#include "common.cuh"
#include "fft_device.cuh"
#include <cmath>
#include <cstdlib>

using namespace clb200;
using namespace clb200::fftdev;

namespace {

// X = [hist (K-1 samples) | in (n_in samples)], zero beyond
__device__ __forceinline__ float2 stream_at(const float2 *__restrict__ hist,
                                            const float2 *__restrict__ in, long p, int km1, long n_in)
{
    if (p < km1) return __ldg(hist + p);
    p -= km1;
    return p < n_in ? __ldg(in + p) : make_float2(0.f, 0.f);
}

// ---------------------------------------------------------------- FFT filter --
template <int LOGN, int EPT, int MINB>
__device__ __forceinline__ void fftfilt_body(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
          float2 *__restrict__ out, const float2 *__restrict__ H, const float2 *__restrict__ tw,
          int K, long nblocks, int D, int skip, long blk0, unsigned long long *wq)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, NPASS = P::npass();
    constexpr int R0 = P::radix(0), RL = P::radix(NPASS - 1), NSL = P::ns(NPASS - 1);
    constexpr int LRL = ilog2(RL);
    // twiddles multiplied up from the power-of-two table entries for the multi-warp block sizes; the one-warp
    // 1024-point blocks are issue-latency bound and lose 4 % with it
    constexpr int TWD = P::T > 32 ? 2 : 1;
    extern __shared__ __align__(16) float2 smem[];
    const int lt = threadIdx.x;
    const int km1 = K - 1;
    const int L = N - km1;

    // blocks come from the work counter when there is one (common.cuh: tile_fetch), else by static striding
    __shared__ long s_next;
    for (long blk = blk0 + blockIdx.x; blk < nblocks;) {
        long nxt = blk + gridDim.x;
        if (wq != nullptr && lt == 0) nxt = blk0 + tile_fetch(wq);
        const long base = blk * (long)L;
        float2 x[EPT];
        // interior blocks (no history, no zero fill): plain streaming loads at
        // compile-time offsets; only the first and last block take the checked path
        const bool interior = base >= km1 && base + N <= n_in + km1;
        if (interior) {
            const float2 *src = in + (base - km1) + lt;
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                x[e] = ldg_stream(src + in_index<P, EPT>(0, e));
            });
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++)
                x[e] = stream_at(hist, in, base + in_index<P, EPT>(lt, e), km1, n_in);
        }

        fft_core<P, EPT, NoHook, TWD>(x, smem, lt, tw);

        // spectrum * H, then re-order into first-pass order with re/im swapped (inverse)
        float2 y[EPT];
        if constexpr (R0 == RL) {
            // output (u, r) is element lt + u*T + r*(N/R): the very slot the next
            // transform's first pass wants -- stay in registers
            static_for<0, EPT / RL>([&](auto u_) {
                constexpr int u = decltype(u_)::value;
                static_for<0, RL>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    const int o = lt + u * P::T + r * NSL;
                    float2 a = cmul(x[u * RL + bitrev(r, LRL)], __ldg(H + o));
                    y[u * RL + r] = make_float2(a.y, a.x);
                });
            });
        } else {
            __syncthreads();      // last pass' shared-memory reads are done
            for_each_output<P, EPT>(x, lt, [&](int o, float2 a) {
                a = cmul(a, __ldg(H + o));
                smem[P::pad(o)] = make_float2(a.y, a.x);
            });
            __syncthreads();
#pragma unroll
            for (int e = 0; e < EPT; e++) y[e] = smem[P::pad(in_index<P, EPT>(lt, e))];
        }

        fft_core<P, EPT, NoHook, TWD>(y, smem, lt, tw);

        if (D == 1 && base + L <= n_in) {
            // every valid output of the block exists: one predicate per element
            float2 *dst = out + (base - km1) + lt;
            for_each_output_c<P, EPT>(y, [&](auto c_, float2 a) {
                constexpr int c = decltype(c_)::value;
                if (c + lt >= km1) __stcs(dst + c, make_float2(a.y, a.x));
            });
        } else {
            for_each_output<P, EPT>(y, lt, [&](int n, float2 a) {
                if (n >= km1) {
                    long m = base + (n - km1);      // output index within this call
                    if (m < n_in) {
                        long q = m - skip;
                        float2 v = make_float2(a.y, a.x);
                        if (D == 1) {
                            __stcs(out + q, v);
                        } else if (q >= 0 && q % D == 0) {
                            __stcs(out + q / D, v);
                        }
                    }
                }
            });
        }
        if (wq != nullptr) {
            if constexpr (P::T == 32) {
                nxt = __shfl_sync(0xffffffffu, nxt, 0);
            } else {
                if (lt == 0) s_next = nxt;
                __syncthreads();         // (the barriers inside fft_core keep the next hand-over behind this read)
                nxt = s_next;
            }
        }
        blk = nxt;
    }
    if (wq != nullptr && lt == 0) tile_finish(wq);
}

template <int LOGN, int EPT, int MINB>
__global__ void __launch_bounds__((1 << LOGN) / EPT, MINB)
k_fftfilt(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
          float2 *__restrict__ out, const float2 *__restrict__ H, const float2 *__restrict__ tw,
          int K, long nblocks, int D, int skip, long blk0, unsigned long long *wq)
{
    fftfilt_body<LOGN, EPT, MINB>(hist, in, n_in, out, H, tw, K, nblocks, D, skip, blk0, wq);
}
// register-capped instantiations of the one-warp kernel (A/B: CLB200_FILT_MINB = 13 / 14 / 15 -> 152 / 144 / 136 registers)
template <int LOGN, int EPT, int NREG>
__global__ void __maxnreg__(NREG)
k_fftfilt_r(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
            float2 *__restrict__ out, const float2 *__restrict__ H, const float2 *__restrict__ tw,
            int K, long nblocks, int D, int skip, long blk0, unsigned long long *wq)
{
    fftfilt_body<LOGN, EPT, 1>(hist, in, n_in, out, H, tw, K, nblocks, D, skip, blk0, wq);
}

// One-warp blocks with the NEXT block's samples fetched ahead of time (interior blocks, 16 B aligned streams):
//   PF = 1: one lane asks the bulk-copy engine to pull the block into L2 (cp.async.bulk.prefetch.L2) -- the loads of the
//           next iteration then see L2 latency instead of a loaded HBM's;
//   PF = 2: one lane starts a bulk copy of the block into a second shared-memory line, completion on an mbarrier;
//           the first pass reads its inputs with LDS.64 (lanes read 256 contiguous bytes: conflict-free).
// The profile of the plain kernel (profiles/r2_fftfilt_full.txt) has 18 % of all warp time on the first butterfly
// waiting for the block's 32 HBM loads (~2600 cycles per block at 3 warps per scheduler).
template <int LOGN, int EPT, int MINB, int PF>
__global__ void __launch_bounds__((1 << LOGN) / EPT, MINB)
k_fftfilt_pf(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
             float2 *__restrict__ out, const float2 *__restrict__ H, const float2 *__restrict__ tw,
             int K, long nblocks, int D, int skip, long blk0, unsigned long long *wq)
{
    using P = Plan<LOGN, EPT>;
    constexpr int N = P::N, NPASS = P::npass();
    constexpr int R0 = P::radix(0), RL = P::radix(NPASS - 1), NSL = P::ns(NPASS - 1);
    constexpr int LRL = ilog2(RL);
    static_assert(P::T == 32 && R0 == RL, "one-warp blocks");
    constexpr uint32_t LINE_BYTES = (N + 2) * (uint32_t)sizeof(float2);
    extern __shared__ __align__(16) float2 smem[];
    float2 *line = smem + P::SMEM_F2;                                   // PF == 2: N + 2 samples
    uint64_t *bar = reinterpret_cast<uint64_t *>(line + N + 2);
    const int lt = threadIdx.x;
    const int km1 = K - 1;
    const int L = N - km1;

    // block b can be fetched ahead when all of [start & ~1, +N+2) lies inside `in`
    auto ahead_ok = [&](long b) {
        const long start = b * (long)L - km1;
        return b < nblocks && start >= 0 && (start & ~1L) + N + 2 <= n_in;
    };
    auto issue = [&](long b) {
        const long start = (b * (long)L - km1) & ~1L;
        if (lt == 0) {
            if constexpr (PF == 2) {
                mbar_expect_tx(bar, LINE_BYTES);
                bulk_g2s(line, in + start, LINE_BYTES, bar);
            } else {
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + start), "r"(LINE_BYTES) : "memory");
            }
        }
    };
    uint32_t parity = 0;
    if constexpr (PF == 2) {
        if (lt == 0) mbar_init(bar, 1);
        __syncwarp();
    }
    if (ahead_ok(blk0 + blockIdx.x)) issue(blk0 + blockIdx.x);

    for (long blk = blk0 + blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long base = blk * (long)L;
        float2 x[EPT];
        const bool ahead = ahead_ok(blk);
        const bool interior = base >= km1 && base + N <= n_in + km1;
        if (PF == 2 && ahead) {
            mbar_wait(bar, parity);
            parity ^= 1;
            const float2 *src = line + ((base - km1) & 1) + lt;
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                x[e] = src[in_index<P, EPT>(0, e)];
            });
            __syncwarp();                                               // every lane has its samples: the line is free
        } else if (interior) {
            const float2 *src = in + (base - km1) + lt;
            static_for<0, EPT>([&](auto e_) {
                constexpr int e = decltype(e_)::value;
                x[e] = ldg_stream(src + in_index<P, EPT>(0, e));
            });
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++)
                x[e] = stream_at(hist, in, base + in_index<P, EPT>(lt, e), km1, n_in);
        }
        if (ahead_ok(blk + gridDim.x)) issue(blk + gridDim.x);

        fft_core<P, EPT, NoHook, 1>(x, smem, lt, tw);

        float2 y[EPT];
        static_for<0, EPT / RL>([&](auto u_) {
            constexpr int u = decltype(u_)::value;
            static_for<0, RL>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const int o = lt + u * P::T + r * NSL;
                float2 a = cmul(x[u * RL + bitrev(r, LRL)], __ldg(H + o));
                y[u * RL + r] = make_float2(a.y, a.x);
            });
        });

        fft_core<P, EPT, NoHook, 1>(y, smem, lt, tw);

        if (D == 1 && base + L <= n_in) {
            float2 *dst = out + (base - km1) + lt;
            for_each_output_c<P, EPT>(y, [&](auto c_, float2 a) {
                constexpr int c = decltype(c_)::value;
                if (c + lt >= km1) __stcs(dst + c, make_float2(a.y, a.x));
            });
        } else {
            for_each_output<P, EPT>(y, lt, [&](int n, float2 a) {
                if (n >= km1) {
                    long m = base + (n - km1);
                    if (m < n_in) {
                        long q = m - skip;
                        float2 v = make_float2(a.y, a.x);
                        if (D == 1) {
                            __stcs(out + q, v);
                        } else if (q >= 0 && q % D == 0) {
                            __stcs(out + q / D, v);
                        }
                    }
                }
            });
        }
    }
}

// ---- the same one-warp 1024-point blocks with HALF the instruction footprint ----------------------------------
// k_fftfilt's loop body is ~2100 straight-line instructions (33 KB) and every one of an SM's 12 warps walks through
// it on its own: ncu shows the SM instruction caches missing to the GPC-level cache at 64 % of THAT cache's request
// peak (gcc__cache_requests_type_instruction; every other kernel of the library: < 1.1 %), and a register-only probe
// of the same butterfly code falls from 34 to 17 T lane-ops/s once its loop body grows from 30 to 57 KB
// (tools/src/bfly_probe.cu).  Here both transforms of a block run through ONE copy of the code: pass 0 is a
// decimation-in-frequency butterfly (natural -> bit-reversed registers), pass 1 a decimation-in-time butterfly
// (bit-reversed -> natural; the exchange through shared memory permutes for free), so the spectrum comes back in the
// very register order the next transform starts from and `for (ph = 0; ph < 2; ph++)` needs no register shuffle.
// Only interior blocks (no history, no ragged end, decimation 1); the caller runs the generic kernel on the edges.
template <int PF>
__global__ void __launch_bounds__(32, 12)
k_fftfilt_1w(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ H,
             const float2 *__restrict__ tw, int K, long blk0, long blk1, unsigned long long *wq)
{
    using P = Plan<10, 32>;
    constexpr int N = 1024, R = 32;
    constexpr uint32_t LINE_BYTES = (N + 2) * (uint32_t)sizeof(float2);
    extern __shared__ __align__(16) float2 smem[];
    float2 *line = smem + P::SMEM_F2;                                   // PF: N + 2 samples of the next block
    uint64_t *bar = reinterpret_cast<uint64_t *>(line + N + 2);
    const int lt = threadIdx.x;
    const int km1 = K - 1;
    const int L = N - km1;
    float2 *const ldp = smem + P::pad(lt);
    float2 *const stp = smem + P::pad(lt * R);
    const float2 *const twp = tw + lt;
    const float2 *const Hp = H + lt;

    auto issue = [&](long b) {
        if (lt == 0) {
            mbar_expect_tx(bar, LINE_BYTES);
            bulk_g2s(line, in + ((b * (long)L - km1) & ~1L), LINE_BYTES, bar);
        }
    };
    uint32_t parity = 0;
    if constexpr (PF) {
        if (lt == 0) mbar_init(bar, 1);
        __syncwarp();
        if (blk0 + blockIdx.x < blk1) issue(blk0 + blockIdx.x);
    }

    long nxt = blk0 + blockIdx.x + gridDim.x;                // PF: the block after the one in flight
    if (PF && wq != nullptr) {
        if (lt == 0) nxt = blk0 + tile_fetch(wq);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
    }
    for (long blk = blk0 + blockIdx.x; blk < blk1;) {
        if (!PF) {
            nxt = blk + gridDim.x;
            if (wq != nullptr && lt == 0) nxt = blk0 + tile_fetch(wq);
        }
        const long start = blk * (long)L - km1;              // first stream sample of the block (>= 0 for these blocks)
        float2 x[R];
        if constexpr (PF) {
            mbar_wait(bar, parity);
            parity ^= 1;
            const float2 *src = line + (start & 1) + lt;
            static_for<0, R>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                x[r] = src[r * 32];
            });
            __syncwarp();                                    // every lane has its samples: the line is free again
            if (nxt < blk1) issue(nxt);
        } else {
            const float2 *src = in + start + lt;
            static_for<0, R>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                x[r] = ldg_stream(src + r * 32);
            });
        }
#pragma unroll 1
        for (int ph = 0; ph < 2; ph++) {
            // pass 0: x[r] = element lt + 32 r  ->  X0[k] in x[bitrev(k)]  ->  shared memory [lt * 32 + k]
            dft_dif<R, 0>(x);
            __syncwarp();                                    // the previous transform's pass-1 reads are done
            static_for<0, R / 2>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                const float2 a = x[bitrev(2 * q, 5)], b = x[bitrev(2 * q + 1, 5)];
                *reinterpret_cast<float4 *>(stp + P::pad(2 * q)) = make_float4(a.x, a.y, b.x, b.y);
            });
            __syncwarp();
            // pass 1: element lt + 32 r, times W_1024^(r * lt), into the slot a decimation-in-time butterfly wants
            static_for<0, R>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const float2 v = ldp[P::pad(r * 32)];
                if constexpr (r == 0) x[0] = v;
                else x[bitrev(r, 5)] = cmul(v, __ldg(twp + (r - 1) * 32));
            });
            dft_dit<R, 0>(x);                                // x[r] = element lt + 32 r of the transform
            if (ph == 0) {
                // spectrum * H; re/im swapped: the second round is the inverse transform (forward on swapped data)
                static_for<0, R>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    const float2 a = cmul(x[r], __ldg(Hp + r * 32));
                    x[r] = make_float2(a.y, a.x);
                });
            }
        }
        float2 *dst = out + start + lt;
        static_for<0, R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            if (r * 32 + lt >= km1) __stcs(dst + r * 32, make_float2(x[r].y, x[r].x));
        });
        if constexpr (PF) {
            // the next block's copy is in flight; fetch the one after it
            blk = nxt;
            nxt = blk + gridDim.x;
            if (wq != nullptr) {
                if (lt == 0) nxt = blk0 + tile_fetch(wq);
                nxt = __shfl_sync(0xffffffffu, nxt, 0);
            }
        } else {
            if (wq != nullptr) nxt = __shfl_sync(0xffffffffu, nxt, 0);
            blk = nxt;
        }
    }
    if (wq != nullptr && lt == 0) tile_finish(wq);
}

// ------------------------------------------------------------ time-domain FIR --
constexpr int FIR_THREADS = 256;
constexpr int FIR_OPT = 8;                         // outputs per thread
constexpr int FIR_TILE = FIR_THREADS * FIR_OPT;    // outputs per CTA

__host__ __device__ constexpr int fir_pad(int i) { return i + (i >> 3); }

// decimation 1: register sliding window, taps (reversed, zero-padded to K8) and
// the input tile in shared memory
// PK: the (re, im) pair of an output is accumulated by ONE packed fma (fma.rn.f32x2 -> FFMA2; the tap sits in both
// halves of a register pair, read as {t, t} from shared memory): 64 instead of 128 fma instructions per step, same bits.
template <bool PK>
__global__ void __launch_bounds__(FIR_THREADS)
k_fir_d1(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
         float2 *__restrict__ out, const float *__restrict__ rtaps, int K, int K8, unsigned long long *wq)
{
    __shared__ long s_next;
    extern __shared__ __align__(16) unsigned char fir_smem[];
    float *s_t = reinterpret_cast<float *>(fir_smem);                     // K8 floats (PK: K8 pairs {t, t})
    float2 *s_t2 = reinterpret_cast<float2 *>(fir_smem);
    float2 *s_x = reinterpret_cast<float2 *>(fir_smem + (size_t)K8 * (PK ? 8 : 4));  // fir_pad(TILE + K8)
    const int km1 = K - 1;
    for (int i = threadIdx.x; i < K8; i += FIR_THREADS) {
        if (PK) s_t2[i] = make_float2(rtaps[i], rtaps[i]);
        else s_t[i] = rtaps[i];
    }
    const long ntile = (n_in + FIR_TILE - 1) / FIR_TILE;
    for (long tile = blockIdx.x; tile < ntile;) {
        long nxt = tile + gridDim.x;                      // tiles from the work counter (common.cuh) when there is one
        if (wq != nullptr && threadIdx.x == 0) nxt = tile_fetch(wq);
        const long g0 = tile * FIR_TILE;
        __syncthreads();
        for (int i = threadIdx.x; i < FIR_TILE + K8; i += FIR_THREADS)
            s_x[fir_pad(i)] = stream_at(hist, in, g0 + i, km1, n_in);
        __syncthreads();
        const int o0 = threadIdx.x * FIR_OPT;
        float2 acc[FIR_OPT], w[2 * FIR_OPT];
        // o0 and the tap index advance in steps of 8, so fir_pad() is linear along the walk: 9 slots per 8 samples.
        // One pointer bump per step and immediate offsets instead of a shift-and-add per load (the loop was 128 FFMA
        // in ~180 instructions); two steps per iteration so that the window hand-over is a renaming, not 16 MOVs.
        const float2 *wp = s_x + fir_pad(o0);
#pragma unroll
        for (int j = 0; j < FIR_OPT; j++) {
            acc[j] = make_float2(0.f, 0.f);
            w[j] = wp[j];
        }
#pragma unroll 2
        for (int i = 0; i < K8; i += FIR_OPT) {
            wp += FIR_OPT + 1;
#pragma unroll
            for (int j = 0; j < FIR_OPT; j++) w[FIR_OPT + j] = wp[j];
            if constexpr (PK) {
                float2 tt[FIR_OPT];
#pragma unroll
                for (int q = 0; q < FIR_OPT / 2; q++) {
                    const float4 t = *reinterpret_cast<const float4 *>(s_t2 + i + 2 * q);
                    tt[2 * q] = make_float2(t.x, t.y);
                    tt[2 * q + 1] = make_float2(t.z, t.w);
                }
#pragma unroll
                for (int ii = 0; ii < FIR_OPT; ii++)
#pragma unroll
                    for (int j = 0; j < FIR_OPT; j++) acc[j] = cfma_real(tt[ii], w[ii + j], acc[j]);
            } else {
                const float4 t0 = *reinterpret_cast<const float4 *>(s_t + i);
                const float4 t1 = *reinterpret_cast<const float4 *>(s_t + i + 4);
                const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int ii = 0; ii < FIR_OPT; ii++)
#pragma unroll
                    for (int j = 0; j < FIR_OPT; j++) {
                        acc[j].x = fmaf(t[ii], w[ii + j].x, acc[j].x);
                        acc[j].y = fmaf(t[ii], w[ii + j].y, acc[j].y);
                    }
            }
#pragma unroll
            for (int j = 0; j < FIR_OPT; j++) w[j] = w[FIR_OPT + j];
        }
#pragma unroll
        for (int j = 0; j < FIR_OPT; j++)
            if (g0 + o0 + j < n_in) __stcs(out + g0 + o0 + j, acc[j]);
        if (wq != nullptr) {
            if (threadIdx.x == 0) s_next = nxt;
            __syncthreads();        // (the two barriers of the next tile's load keep its hand-over behind this read)
            nxt = s_next;
        }
        tile = nxt;
    }
    if (wq != nullptr && threadIdx.x == 0) tile_finish(wq);
}

// any decimation: one output per thread, taps in shared memory
__global__ void __launch_bounds__(FIR_THREADS)
k_fir_dec(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
          float2 *__restrict__ out, long n_out, const float *__restrict__ rtaps, int K, int D,
          int skip)
{
    extern __shared__ __align__(16) unsigned char fir_smem[];
    float *s_t = reinterpret_cast<float *>(fir_smem);
    for (int i = threadIdx.x; i < K; i += FIR_THREADS) s_t[i] = rtaps[i];
    __syncthreads();
    const int km1 = K - 1;
    const long stride = (long)gridDim.x * FIR_THREADS;
    for (long o = (long)blockIdx.x * FIR_THREADS + threadIdx.x; o < n_out; o += stride) {
        const long g = skip + o * D;
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < K; i++) {
            float2 v = stream_at(hist, in, g + i, km1, n_in);
            acc.x = fmaf(s_t[i], v.x, acc.x);
            acc.y = fmaf(s_t[i], v.y, acc.y);
        }
        out[o] = acc;
    }
}

// new history = last K-1 samples of [hist | in]
__global__ void k_hist_update(const float2 *__restrict__ hist, const float2 *__restrict__ in,
                              long n_in, float2 *__restrict__ nhist, int km1)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < km1; i += gridDim.x * blockDim.x) {
        long p = n_in + i;      // position in [hist | in] of new history sample i
        nhist[i] = p < km1 ? hist[p] : in[p - km1];
    }
}

// ---------------------------------------------------------------------- host --
typedef void (*fftfilt_kernel_t)(const float2 *, const float2 *, long, float2 *, const float2 *,
                                 const float2 *, int, long, int, int, long, unsigned long long *);

struct FiltVariant {
    int logn, threads, smem_bytes;
    void (*fill_tw)(std::vector<float2> &);
    fftfilt_kernel_t kernel;
    fftfilt_kernel_t kernel_pf[2] = {nullptr, nullptr};   // one-warp blocks: next block prefetched to L2 / into shared memory
    int smem_pf[2] = {0, 0};
};

template <int LOGN, int EPT>
void fill_tw_f(std::vector<float2> &tw)
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

template <int LOGN, int EPT, int MINB>
FiltVariant make_filt()
{
    using P = Plan<LOGN, EPT>;
    FiltVariant v{LOGN, P::T, P::SMEM_F2 * (int)sizeof(float2), &fill_tw_f<LOGN, EPT>, &k_fftfilt<LOGN, EPT, MINB>};
    if constexpr (P::T == 32 && P::npass() == 2) {
        v.kernel_pf[0] = &k_fftfilt_pf<LOGN, EPT, MINB, 1>;
        v.kernel_pf[1] = &k_fftfilt_pf<LOGN, EPT, MINB, 2>;
        v.smem_pf[0] = v.smem_bytes;
        v.smem_pf[1] = v.smem_bytes + (P::N + 2) * (int)sizeof(float2) + 16;
    }
    return v;
}

// block FFT size by tap count: keep L = NF-K+1 >= NF/2 so at most half of every
// transform is overlap
const FiltVariant *pick_filt(int ntaps)
{
    static const FiltVariant v10 = make_filt<10, 32, 12>();   // 1024 = 32^2: one warp per block, 1 exchange per FFT
                                                              // (12 CTAs/SM, 170 regs: 3.15 TB/s vs 2.76 for 4096 at 256 taps)
    static const FiltVariant v10b = make_filt<10, 32, 16>();  // same, capped at 128 registers (A/B: CLB200_FILT_MINB=16)
    static FiltVariant v10c = make_filt<10, 32, 12>(), v10d = make_filt<10, 32, 12>(), v10e = make_filt<10, 32, 12>();
    v10c.kernel = &k_fftfilt_r<10, 32, 152>;
    v10d.kernel = &k_fftfilt_r<10, 32, 144>;
    v10e.kernel = &k_fftfilt_r<10, 32, 136>;
    static const FiltVariant v12 = make_filt<12, 16, 3>();    // 4096 = 16^3 (register hand-over), three CTAs per SM (+3 % over
                                                              // two; 32 x 32 x 4 with a shared-memory re-order measured slower)
    static const FiltVariant v14 = make_filt<14, 32, 1>();    // 16384 = 32 * 32 * 16, 512 threads (16^3 * 4 with 1024 threads and
                                                              // 64 registers measured 18 % slower: 1.38 vs 1.69 TB/s at 3000 taps)
    const char *e = getenv("CLB200_FILT_NF");                 // tuning: force the block size
    const int force = e ? atoi(e) : 0;
    if (force == 1024 && ntaps <= 1024) return &v10;
    if (force == 4096 && ntaps <= 4096) return &v12;
    if (ntaps <= 513 && force == 0) {
        const char *mb = getenv("CLB200_FILT_MINB");
        const int m = mb ? atoi(mb) : 0;
        return m == 16 ? &v10b : m == 15 ? &v10e : m == 14 ? &v10d : m == 13 ? &v10c : &v10;
    }
    if (ntaps <= 2049) return &v12;
    if (ntaps <= 8193) return &v14;
    return nullptr;
}

struct Filter : clb200_block {
    int decim = 1, use_time = 0;
    std::vector<float> taps, pending;
    bool updated = false;
    bool time_kernel = false;           // which kernel runs (FFT mode falls back for huge K)
    const FiltVariant *var = nullptr;
    Buf d_hist[2], d_H, d_tw, d_rtaps;
    int cur = 0;                        // which d_hist is live
    int skip = 0;                       // decimation phase: inputs to drop before the next output
    int k8 = 0;
    int resident = 1;
    bool fir_packed = true;             // time-domain kernel with packed fma (k_fir_d1<true>)
    int pf = 0, resident_pf = 1;        // prefetching one-warp kernel (0: off, 1: L2, 2: shared memory) and its occupancy
    int compact = 0, resident_c = 1, smem_c = 0;   // compact one-warp kernel k_fftfilt_1w (1: plain loads, 2: bulk-copy prefetch)
    cudaEvent_t hist_ready = nullptr;
    bool hist_pending = false;
    ~Filter() override
    {
        DeviceGuard g(device);
        for (int i = 0; i < 2; i++) d_hist[i].release();
        d_H.release();
        d_tw.release();
        d_rtaps.release();
        if (hist_ready) cudaEventDestroy(hist_ready);
    }
};

// resident CTAs per SM of the time-domain kernel (tuning: CLB200_FIR_CTAS)
int fir_ctas_per_sm()
{
    static const int v = [] {
        const char *e = getenv("CLB200_FIR_CTAS");
        const int n = e ? atoi(e) : 0;
        return n >= 1 && n <= 8 ? n : 0;          // 0: what fits
    }();
    return v;
}

// (re)build everything that depends on the taps; resets the stream state
int filter_configure(Filter *f, const std::vector<float> &taps)
{
    const int K = (int)taps.size();
    f->taps = taps;
    f->skip = 0;
    f->cur = 0;
    f->var = f->use_time ? nullptr : pick_filt(K);
    f->time_kernel = (f->var == nullptr);
    CLB_CUDA(cudaDeviceSynchronize());
    size_t hb = sizeof(float2) * (size_t)std::max(1, K - 1);
    for (int i = 0; i < 2; i++) {
        CLB_TRY(f->d_hist[i].reserve(hb));
        CLB_CUDA(cudaMemset(f->d_hist[i].p, 0, hb));
    }
    if (f->time_kernel) {
        f->k8 = (K + 7) & ~7;
        std::vector<float> rt(f->k8, 0.f);
        for (int i = 0; i < K; i++) rt[i] = taps[K - 1 - i];      // FilterArray[K-1-i] (:187)
        CLB_TRY(f->d_rtaps.reserve(sizeof(float) * f->k8));
        CLB_CUDA(cudaMemcpy(f->d_rtaps.p, rt.data(), sizeof(float) * f->k8, cudaMemcpyHostToDevice));
        {
            const char *pk = getenv("CLB200_FIR_PACKED");         // A/B
            f->fir_packed = pk ? atoi(pk) != 0 : true;
        }
        size_t smem = (size_t)f->k8 * (f->fir_packed ? 8 : 4) + sizeof(float2) * fir_pad(FIR_TILE + f->k8 + 8);
        CLB_CHECK(smem <= 200 * 1024, CLB200_EINVAL, "clFilter: %d taps exceed the FIR kernel's shared memory", K);
        // the limit is per-function state shared by every clFilter handle of the process: always the cap the
        // kernels may need, never the size of the filter configured last
        CLB_CUDA(cudaFuncSetAttribute((const void *)k_fir_d1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CLB_CUDA(cudaFuncSetAttribute((const void *)k_fir_d1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CLB_CUDA(cudaFuncSetAttribute((const void *)k_fir_dec, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // as many resident CTAs as fit (5 at 256 taps): one CTA's tile load and barriers hide behind the
        // others' FMA loops -- 32.2 (2 CTAs/SM) -> 38.0 Gsamples/s at 256 taps
        int occ = 0;
        CLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, f->fir_packed ? (const void *)k_fir_d1<true> : (const void *)k_fir_d1<false>, FIR_THREADS, smem));
        f->resident = std::max(1, std::min(occ, 8));
        if (fir_ctas_per_sm() > 0) f->resident = fir_ctas_per_sm();
    } else {
        const int NF = 1 << f->var->logn;
        // H[k] = (1/NF) * sum_n taps[n] e^{-2 pi i n k / NF}   (fft_filter.cc:52-63)
        std::vector<double> cs(NF), sn(NF);
        for (int i = 0; i < NF; i++) {
            cs[i] = cos(-2.0 * M_PI * i / NF);
            sn[i] = sin(-2.0 * M_PI * i / NF);
        }
        std::vector<float2> H(NF);
        for (int k = 0; k < NF; k++) {
            double re = 0, im = 0;
            for (int n = 0; n < K; n++) {
                int idx = (int)(((long)n * k) & (NF - 1));
                re += taps[n] * cs[idx];
                im += taps[n] * sn[idx];
            }
            H[k] = make_float2((float)(re / NF), (float)(im / NF));
        }
        CLB_TRY(f->d_H.reserve(sizeof(float2) * NF));
        CLB_CUDA(cudaMemcpy(f->d_H.p, H.data(), sizeof(float2) * NF, cudaMemcpyHostToDevice));
        std::vector<float2> tw;
        f->var->fill_tw(tw);
        CLB_TRY(f->d_tw.reserve(sizeof(float2) * tw.size()));
        CLB_CUDA(cudaMemcpy(f->d_tw.p, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
        CLB_CUDA(cudaFuncSetAttribute((const void *)f->var->kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, f->var->smem_bytes));
        int occ = 0;
        CLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)f->var->kernel,
                                                               f->var->threads, f->var->smem_bytes));
        CLB_CHECK(occ >= 1, CLB200_ECUDA, "clFilter: FFT-filter kernel does not fit an SM");
        f->resident = occ;
        f->compact = 0;
        if (f->var->logn == 10 && f->var->threads == 32) {
            const char *ce = getenv("CLB200_FILT_COMPACT");
            const int wantc = ce ? atoi(ce) : 0;      // opt-in (measured on par with the generic kernel)
            if (wantc == 1 || wantc == 2) {
                const void *k = wantc == 2 ? (const void *)k_fftfilt_1w<1> : (const void *)k_fftfilt_1w<0>;
                const int sm = f->var->smem_bytes + (wantc == 2 ? (1024 + 2) * (int)sizeof(float2) + 16 : 0);
                int occ2 = 0;
                if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k, 32, sm) == cudaSuccess && occ2 >= 1) {
                    f->compact = wantc;
                    f->resident_c = occ2;
                    f->smem_c = sm;
                }
                cudaGetLastError();
            }
        }
        f->pf = 0;
        const char *pe = getenv("CLB200_FILT_PF");
        const int want = pe ? atoi(pe) : 0;
        if ((want == 1 || want == 2) && f->var->kernel_pf[want - 1]) {
            const void *k = (const void *)f->var->kernel_pf[want - 1];
            int occ2 = 0;
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, f->var->smem_pf[want - 1]) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k, f->var->threads, f->var->smem_pf[want - 1]) == cudaSuccess &&
                occ2 >= 1) {
                f->pf = want;
                f->resident_pf = occ2;
            }
            cudaGetLastError();
        }
    }
    if (f->time_kernel)
        f->set_info("clFilter %d taps, decimation %d: time-domain k_fir_d1 (register sliding window, taps + %d-sample tile in shared "
                    "memory), %d threads, %d CTAs/SM", K, f->decim, FIR_TILE, FIR_THREADS, f->resident);
    else
        f->set_info("clFilter %d taps, decimation %d: overlap-save k_fftfilt, %d-pt blocks (%d new outputs each; the reference blocks "
                    "%d/%d), %d threads, %d B shared memory, %d CTAs/SM", K, f->decim, 1 << f->var->logn, (1 << f->var->logn) - K + 1,
                    (int)(2 * pow(2.0, ceil(log((double)K) / log(2.0)))), (int)(2 * pow(2.0, ceil(log((double)K) / log(2.0)))) - K + 1,
                    f->var->threads, f->var->smem_bytes, f->resident);
    return CLB200_OK;
}

int apply_pending(Filter *f)
{
    std::vector<float> t;
    {
        std::lock_guard<std::mutex> g(f->mtx);
        if (!f->updated) return CLB200_OK;
        t = f->pending;
        f->updated = false;
    }
    return filter_configure(f, t);
}

// enqueue one chunk on `st`; d_in holds n_in new samples
int filter_launch(Filter *f, const float2 *d_in, long n_in, float2 *d_out, long *n_out,
                  cudaStream_t st)
{
    const int K = (int)f->taps.size(), D = f->decim, km1 = K - 1;
    long nout = n_in > f->skip ? (n_in - f->skip + D - 1) / D : 0;
    if (n_out) *n_out = nout;
    if (n_in <= 0) return CLB200_OK;
    const int sms = device_sm_count(f->device);
    if (f->hist_pending) CLB_CUDA(cudaStreamWaitEvent(st, f->hist_ready, 0));
    const float2 *hist = (const float2 *)f->d_hist[f->cur].p;
    float2 *nhist = (float2 *)f->d_hist[f->cur ^ 1].p;
    if (nout > 0) {
        if (f->time_kernel) {
            if (D == 1) {
                long ntile = (n_in + FIR_TILE - 1) / FIR_TILE;
                size_t smem = (size_t)f->k8 * (f->fir_packed ? 8 : 4) + sizeof(float2) * fir_pad(FIR_TILE + f->k8 + 8);
                (f->fir_packed ? k_fir_d1<true> : k_fir_d1<false>)<<<grid_for(ntile, sms, f->resident), FIR_THREADS, smem, st>>>(
                    hist, d_in, n_in, d_out, (const float *)f->d_rtaps.p, K, f->k8,
                    ntile > (long)sms * f->resident ? f->work_counter(st) : nullptr);
            } else {
                long ctas = (nout + FIR_THREADS - 1) / FIR_THREADS;
                k_fir_dec<<<grid_for(ctas, sms, 8), FIR_THREADS, f->k8 * 4, st>>>(
                    hist, d_in, n_in, d_out, nout, (const float *)f->d_rtaps.p, K, D, f->skip);
            }
        } else {
            const int NF = 1 << f->var->logn, L = NF - km1;
            long nblocks = (n_in + L - 1) / L;
            const float2 *dH = (const float2 *)f->d_H.p, *dtw = (const float2 *)f->d_tw.p;
            // one-warp 1024-point blocks, decimation 1: the compact kernel takes the interior blocks
            //   b*L - (K-1) >= 0  and  b*L - (K-1) + NF (+ 2 with the bulk-copy prefetch) <= n_in,
            // the generic kernel the (at most three) blocks at the two ends of the call
            long b0 = 0, b1 = 0;
            unsigned long long *wq = nullptr;
            if (f->compact && D == 1) {
                const int slack = f->compact == 2 ? 2 : 0;
                b0 = (km1 + L - 1) / L;
                b1 = n_in - NF - slack + km1 >= 0 ? (n_in - NF - slack + km1) / L + 1 : 0;
                b1 = std::min(b1, nblocks);
                if (f->compact == 2 && ((uintptr_t)d_in & 15) != 0) b1 = b0;      // bulk copies need 16 B aligned streams
            }
            if (b1 > b0) {
                if (b1 - b0 > (long)sms * f->resident_c) wq = f->work_counter(st);
                if (f->compact == 2)
                    k_fftfilt_1w<1><<<grid_for(b1 - b0, sms, f->resident_c), 32, f->smem_c, st>>>(d_in, d_out, dH, dtw, K, b0, b1, wq);
                else
                    k_fftfilt_1w<0><<<grid_for(b1 - b0, sms, f->resident_c), 32, f->smem_c, st>>>(d_in, d_out, dH, dtw, K, b0, b1, wq);
                f->n_launch++;
                if (b0 > 0)
                    f->var->kernel<<<grid_for(b0, sms, f->resident), f->var->threads, f->var->smem_bytes, st>>>(
                        hist, d_in, n_in, d_out, dH, dtw, K, b0, D, f->skip, 0, nullptr);
                if (b1 < nblocks)
                    f->var->kernel<<<grid_for(nblocks - b1, sms, f->resident), f->var->threads, f->var->smem_bytes, st>>>(
                        hist, d_in, n_in, d_out, dH, dtw, K, nblocks, D, f->skip, b1, nullptr);
            } else if (f->pf && ((uintptr_t)d_in & 15) == 0)
                f->var->kernel_pf[f->pf - 1]<<<grid_for(nblocks, sms, f->resident_pf), f->var->threads,
                                               f->var->smem_pf[f->pf - 1], st>>>(hist, d_in, n_in, d_out, dH, dtw, K, nblocks, D, f->skip, 0, nullptr);
            else
                f->var->kernel<<<grid_for(nblocks, sms, f->resident), f->var->threads, f->var->smem_bytes, st>>>(
                    hist, d_in, n_in, d_out, dH, dtw, K, nblocks, D, f->skip, 0,
                    nblocks > (long)sms * f->resident ? f->work_counter(st) : nullptr);
        }
        CLB_CUDA(cudaGetLastError());
        f->n_launch++;
    }
    if (km1 > 0) {
        k_hist_update<<<(km1 + 255) / 256, 256, 0, st>>>(hist, d_in, n_in, nhist, km1);
        CLB_CUDA(cudaGetLastError());
        f->n_launch++;
        f->cur ^= 1;
        if (!f->hist_ready) CLB_CUDA(cudaEventCreateWithFlags(&f->hist_ready, cudaEventDisableTiming));
        CLB_CUDA(cudaEventRecord(f->hist_ready, st));
        f->hist_pending = true;
    }
    f->skip = (int)(f->skip + nout * D - n_in);
    return CLB200_OK;
}

} // namespace

extern "C" {

int clb200_filter_ref_sizes(int ntaps, int *fftsize, int *nsamples)
{
    CLB_CHECK(ntaps >= 1 && fftsize && nsamples, CLB200_EINVAL, "bad arguments");
    // fft_filter.cc:77-78
    int fs = (int)(2 * pow(2.0, ceil(log((double)ntaps) / log(2.0))));
    *fftsize = fs;
    *nsamples = fs - ntaps + 1;
    return CLB200_OK;
}

int clb200_filter_create(int device, int decimation, const float *taps, int ntaps, int use_time,
                         clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(decimation >= 1, CLB200_EINVAL, "clFilter: decimation must be >= 1, got %d", decimation);
    CLB_CHECK(taps != nullptr && ntaps >= 1, CLB200_EINVAL, "clFilter: at least one tap is required");
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    DeviceGuard g(device);
    Filter *f = new Filter;
    f->kind = KIND_FILTER;
    f->device = device;
    f->init_work_counters();
    f->decim = decimation;
    f->use_time = use_time ? 1 : 0;
    int rc = filter_configure(f, std::vector<float>(taps, taps + ntaps));
    if (rc != CLB200_OK) {
        delete f;
        return rc;
    }
    *out = f;
    return CLB200_OK;
}

int clb200_filter_set_taps(clb200_handle h, const float *taps, int ntaps)
{
    Filter *f;
    CLB_TRY(check_kind(h, KIND_FILTER, &f));
    CLB_CHECK(taps != nullptr && ntaps >= 1, CLB200_EINVAL, "clFilter: at least one tap is required");
    std::lock_guard<std::mutex> g(f->mtx);
    f->pending.assign(taps, taps + ntaps);
    f->updated = true;
    return CLB200_OK;
}

int clb200_filter_ntaps(clb200_handle h)
{
    Filter *f;
    if (check_kind(h, KIND_FILTER, &f) != CLB200_OK) return CLB200_EINVAL;
    std::lock_guard<std::mutex> g(f->mtx);
    return (int)(f->updated ? f->pending.size() : f->taps.size());
}

int clb200_filter_get_taps(clb200_handle h, float *taps, int cap)
{
    Filter *f;
    CLB_TRY(check_kind(h, KIND_FILTER, &f));
    std::lock_guard<std::mutex> g(f->mtx);
    const std::vector<float> &t = f->updated ? f->pending : f->taps;
    CLB_CHECK(taps != nullptr && cap >= (int)t.size(), CLB200_EINVAL, "tap buffer too small");
    memcpy(taps, t.data(), sizeof(float) * t.size());
    return CLB200_OK;
}

int clb200_filter_reset(clb200_handle h)
{
    Filter *f;
    CLB_TRY(check_kind(h, KIND_FILTER, &f));
    DeviceGuard g(f->device);
    CLB_TRY(apply_pending(f));
    return filter_configure(f, std::vector<float>(f->taps));
}

int clb200_filter_launch_device(clb200_handle h, const void *d_in, long n_in, void *d_out,
                                long *n_out, void *stream)
{
    Filter *f;
    CLB_TRY(check_kind(h, KIND_FILTER, &f));
    CLB_CHECK(n_in >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(f->device);
    CLB_TRY(apply_pending(f));
    return filter_launch(f, (const float2 *)d_in, n_in, (float2 *)d_out, n_out, (cudaStream_t)stream);
}

int clb200_filter_work(clb200_handle h, const void *in, long n_in, void *out, long *n_out)
{
    Filter *f;
    CLB_TRY(check_kind(h, KIND_FILTER, &f));
    CLB_CHECK(n_in >= 0, CLB200_EINVAL, "negative item count");
    if (n_out) *n_out = 0;
    DeviceGuard g(f->device);
    CLB_TRY(apply_pending(f));
    if (n_in == 0) return CLB200_OK;
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = pd.out_bytes[0] = 8;
    return run_chunked(f, pd, n_in, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *no) {
                           return filter_launch(f, (const float2 *)di[0], n, (float2 *)dout[0], no, st);
                       },
                       n_out);
}

} // extern "C"
