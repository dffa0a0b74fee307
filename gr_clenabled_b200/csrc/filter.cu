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
__global__ void __launch_bounds__((1 << LOGN) / EPT, MINB)
k_fftfilt(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
          float2 *__restrict__ out, const float2 *__restrict__ H, const float2 *__restrict__ tw,
          int K, long nblocks, int D, int skip)
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

    for (long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
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
    }
}

// ------------------------------------------------------------ time-domain FIR --
constexpr int FIR_THREADS = 256;
constexpr int FIR_OPT = 8;                         // outputs per thread
constexpr int FIR_TILE = FIR_THREADS * FIR_OPT;    // outputs per CTA

__host__ __device__ constexpr int fir_pad(int i) { return i + (i >> 3); }

// decimation 1: register sliding window, taps (reversed, zero-padded to K8) and
// the input tile in shared memory
__global__ void __launch_bounds__(FIR_THREADS)
k_fir_d1(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
         float2 *__restrict__ out, const float *__restrict__ rtaps, int K, int K8)
{
    extern __shared__ __align__(16) unsigned char fir_smem[];
    float *s_t = reinterpret_cast<float *>(fir_smem);                     // K8 floats
    float2 *s_x = reinterpret_cast<float2 *>(fir_smem + (size_t)K8 * 4);  // fir_pad(TILE + K8)
    const int km1 = K - 1;
    for (int i = threadIdx.x; i < K8; i += FIR_THREADS) s_t[i] = rtaps[i];
    const long ntile = (n_in + FIR_TILE - 1) / FIR_TILE;
    for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long g0 = tile * FIR_TILE;
        __syncthreads();
        for (int i = threadIdx.x; i < FIR_TILE + K8; i += FIR_THREADS)
            s_x[fir_pad(i)] = stream_at(hist, in, g0 + i, km1, n_in);
        __syncthreads();
        const int o0 = threadIdx.x * FIR_OPT;
        float2 acc[FIR_OPT], w[2 * FIR_OPT];
#pragma unroll
        for (int j = 0; j < FIR_OPT; j++) {
            acc[j] = make_float2(0.f, 0.f);
            w[j] = s_x[fir_pad(o0 + j)];
        }
        for (int i = 0; i < K8; i += FIR_OPT) {
#pragma unroll
            for (int j = 0; j < FIR_OPT; j++) w[FIR_OPT + j] = s_x[fir_pad(o0 + i + FIR_OPT + j)];
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
#pragma unroll
            for (int j = 0; j < FIR_OPT; j++) w[j] = w[FIR_OPT + j];
        }
#pragma unroll
        for (int j = 0; j < FIR_OPT; j++)
            if (g0 + o0 + j < n_in) __stcs(out + g0 + o0 + j, acc[j]);
    }
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
                                 const float2 *, int, long, int, int);

struct FiltVariant {
    int logn, threads, smem_bytes;
    void (*fill_tw)(std::vector<float2> &);
    fftfilt_kernel_t kernel;
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
    return FiltVariant{LOGN, P::T, P::SMEM_F2 * (int)sizeof(float2), &fill_tw_f<LOGN, EPT>,
                       &k_fftfilt<LOGN, EPT, MINB>};
}

// block FFT size by tap count: keep L = NF-K+1 >= NF/2 so at most half of every
// transform is overlap
const FiltVariant *pick_filt(int ntaps)
{
    static const FiltVariant v10 = make_filt<10, 32, 12>();   // 1024 = 32^2: one warp per block, 1 exchange per FFT
                                                              // (12 CTAs/SM, 170 regs: 3.15 TB/s vs 2.76 for 4096 at 256 taps)
    static const FiltVariant v10b = make_filt<10, 32, 16>();  // same, capped at 128 registers (A/B: CLB200_FILT_MINB=16)
    static const FiltVariant v12 = make_filt<12, 16, 2>();    // 4096 = 16^3 (register hand-over)
    static const FiltVariant v14 = make_filt<14, 16, 1>();    // 16384 = 16^3 * 4
    const char *e = getenv("CLB200_FILT_NF");                 // tuning: force the block size
    const int force = e ? atoi(e) : 0;
    if (force == 1024 && ntaps <= 1024) return &v10;
    if (force == 4096 && ntaps <= 4096) return &v12;
    if (ntaps <= 513 && force == 0) {
        const char *mb = getenv("CLB200_FILT_MINB");
        return (mb && atoi(mb) == 16) ? &v10b : &v10;
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
        size_t smem = (size_t)f->k8 * 4 + sizeof(float2) * fir_pad(FIR_TILE + f->k8 + 8);
        CLB_CHECK(smem <= 200 * 1024, CLB200_EINVAL, "clFilter: %d taps exceed the FIR kernel's shared memory", K);
        // the limit is per-function state shared by every clFilter handle of the process: always the cap the
        // kernels may need, never the size of the filter configured last
        CLB_CUDA(cudaFuncSetAttribute((const void *)k_fir_d1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CLB_CUDA(cudaFuncSetAttribute((const void *)k_fir_dec, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // as many resident CTAs as fit (5 at 256 taps): one CTA's tile load and barriers hide behind the
        // others' FMA loops -- 32.2 (2 CTAs/SM) -> 38.0 Gsamples/s at 256 taps
        int occ = 0;
        CLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)k_fir_d1, FIR_THREADS, smem));
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
                size_t smem = (size_t)f->k8 * 4 + sizeof(float2) * fir_pad(FIR_TILE + f->k8 + 8);
                k_fir_d1<<<grid_for(ntile, sms, f->resident), FIR_THREADS, smem, st>>>(
                    hist, d_in, n_in, d_out, (const float *)f->d_rtaps.p, K, f->k8);
            } else {
                long ctas = (nout + FIR_THREADS - 1) / FIR_THREADS;
                k_fir_dec<<<grid_for(ctas, sms, 8), FIR_THREADS, f->k8 * 4, st>>>(
                    hist, d_in, n_in, d_out, nout, (const float *)f->d_rtaps.p, K, D, f->skip);
            }
        } else {
            const int NF = 1 << f->var->logn, L = NF - km1;
            long nblocks = (n_in + L - 1) / L;
            f->var->kernel<<<grid_for(nblocks, sms, f->resident), f->var->threads,
                             f->var->smem_bytes, st>>>(hist, d_in, n_in, d_out,
                                                      (const float2 *)f->d_H.p,
                                                      (const float2 *)f->d_tw.p, K, nblocks, D,
                                                      f->skip);
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
