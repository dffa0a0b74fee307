// extras.cu -- the blocks either side of the hot path (SURVEY 8(f) "next" rows):
//   clXCorrelate (time-domain normalised cross-correlation + arg-max), clComplexFilter
//   (complex-tap FIR), clQuadratureDemod, clSignalSource.
// Each kernel cites the reference kernel string it replaces; the arithmetic order the
// reference prescribes is kept where it defines the result (index/lag outputs, tie rules).
#include "common.cuh"
#include <cmath>
#include <cstdlib>

using namespace clb200;

namespace {

// ================================================================ clXCorrelate ==
// Reference: per frame 2 H2D copies, ComplexToMag x2, F32Squared x2, XCorrelate (one
// work-item per shift, each walking both buffers and both squared buffers sequentially),
// find_max + 2 blocking read-backs, per non-reference input (lib/clXCorrelate_impl.cc:1551-1583).
// Here: one preparation launch for all inputs (magnitudes + fp64 prefix sums of the squares, so
// the two energy sums of a shift are two subtractions instead of two O(L) loops), one launch
// for all shifts of all inputs (a warp per shift, fp64 accumulation), one arg-max launch.
constexpr int XC_GROUP = 1024;      // find_max work-group size of the reference on a 1024-thread device

// mag[k][i] (float) and pre[k][i] = sum_{j<i} mag[k][j]^2 (double, L+1 entries); one CTA per input
__global__ void __launch_bounds__(1024)
k_xc_prep(const void *__restrict__ in, int L, int is_complex, float *__restrict__ mag, double *__restrict__ pre)
{
    __shared__ double s_part[1024];
    const int k = blockIdx.x, t = threadIdx.x;
    const int C = (L + 1023) / 1024;
    const int i0 = min(L, t * C), i1 = min(L, i0 + C);
    float *m = mag + (size_t)k * L;
    double *p = pre + (size_t)k * (L + 1);
    double sum = 0.0;
    for (int i = i0; i < i1; i++) {
        float v;
        if (is_complex) {
            const float2 a = reinterpret_cast<const float2 *>(in)[(size_t)k * L + i];
            v = sqrtf(fmaf(a.x, a.x, a.y * a.y));              // ComplexToMag, fma form (:921)
        } else {
            v = reinterpret_cast<const float *>(in)[(size_t)k * L + i];
        }
        m[i] = v;
        sum += (double)(v * v);                                 // F32Squared (:978) is a float product
    }
    s_part[t] = sum;
    __syncthreads();
    // exclusive scan of the 1024 partial sums (Hillis-Steele; runs once per frame)
    for (int d = 1; d < 1024; d <<= 1) {
        const double add = (t >= d) ? s_part[t - d] : 0.0;
        __syncthreads();
        s_part[t] += add;
        __syncthreads();
    }
    double run = s_part[t] - sum;
    for (int i = i0; i < i1; i++) {
        p[i] = run;
        const float v = m[i];
        run += (double)(v * v);
    }
    if (i1 == L && i0 < L) p[L] = run;
    if (L == 0 && t == 0) p[0] = 0.0;
}

// correlation_factors[k-1][g], g in [0, 2*max_shift): kernel XCorrelate (:851-900)
__global__ void __launch_bounds__(256)
k_xc_factors(const float *__restrict__ mag, const double *__restrict__ pre, int L, int max_shift,
             float *__restrict__ factors)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + warp;
    const int k = blockIdx.y + 1;
    if (g >= 2 * max_shift) return;
    const int shift = g - max_shift;
    const int rs = shift >= 0 ? shift : -shift;
    const int n = L - rs;
    const float *ref = mag + (shift > 0 ? rs : 0);
    const float *sig = mag + (size_t)k * L + (shift > 0 ? 0 : rs);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int i = lane;
    for (; i + 96 < n; i += 128) {
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = fma((double)__ldg(ref + i + 32 * u), (double)__ldg(sig + i + 32 * u), acc[u]);
    }
    for (; i < n; i += 32) acc[0] = fma((double)__ldg(ref + i), (double)__ldg(sig + i), acc[0]);
    double xy = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xy += __shfl_xor_sync(0xffffffffu, xy, o);
    if (lane == 0) {
        float c = -2.0f;
        if (n > 0) {
            const double *pr = pre, *ps = pre + (size_t)k * (L + 1);
            const double x2 = shift > 0 ? pr[L] - pr[rs] : pr[n];
            const double y2 = shift > 0 ? ps[n] : ps[L] - ps[rs];
            const float fx = (float)x2, fy = (float)y2;
            const float denom = fx * fy;
            if (denom != 0.0f) c = (float)xy / sqrtf(fx * fy);
        }
        factors[(size_t)(k - 1) * 2 * max_shift + g] = c;
    }
}

// find_max (:1016-1043): groups of `group` consecutive factors, halving tree, slot id takes slot
// id+stride only when STRICTLY greater; then (host pass :1371-1413) the first strictly greatest
// group.  One CTA per input does both passes; group <= 1024 = blockDim.
__global__ void __launch_bounds__(XC_GROUP)
k_xc_findmax(const float *__restrict__ factors, int n, int group, int max_shift, float *__restrict__ corr,
             int *__restrict__ lag)
{
    __shared__ float s_m[XC_GROUP];
    __shared__ int s_l[XC_GROUP];
    const float *f = factors + (size_t)blockIdx.x * n;
    const int t = threadIdx.x;
    float best = 0.0f;
    int best_i = 0;
    for (int g0 = 0; g0 < n; g0 += group) {
        if (t < group) {
            s_m[t] = f[g0 + t];
            s_l[t] = g0 + t;
        }
        __syncthreads();
        for (int stride = group >> 1; stride > 0; stride >>= 1) {
            if (t < stride && s_m[t + stride] > s_m[t]) {
                s_m[t] = s_m[t + stride];
                s_l[t] = s_l[t + stride];
            }
            __syncthreads();
        }
        if (t == 0 && (g0 == 0 || s_m[0] > best)) {
            best = s_m[0];
            best_i = s_l[0];
        }
        __syncthreads();
    }
    if (t == 0) {
        corr[blockIdx.x] = best;
        lag[blockIdx.x] = best_i - max_shift;
    }
}

struct XCorr : clb200_block {
    int num_inputs = 0, L = 0, dtype = 0, max_shift = 0;
    Buf d_in, d_mag, d_pre, d_fac, d_res, pin_res;
    cudaStream_t st = nullptr;
    bool have_factors = false;
    size_t item() const { return dtype == CLB200_DTYPE_COMPLEX ? 8 : 4; }
    ~XCorr() override
    {
        DeviceGuard g(device);
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        d_in.release();
        d_mag.release();
        d_pre.release();
        d_fac.release();
        d_res.release();
        pin_res.release();
    }
};

int xc_launch(XCorr *x, const void *d_in, float *d_corr, int *d_lag, cudaStream_t st)
{
    const int ns = x->num_inputs - 1, n = 2 * x->max_shift;
    CLB_TRY(x->d_mag.reserve(sizeof(float) * (size_t)x->num_inputs * x->L));
    CLB_TRY(x->d_pre.reserve(sizeof(double) * (size_t)x->num_inputs * (x->L + 1)));
    CLB_TRY(x->d_fac.reserve(sizeof(float) * (size_t)ns * n));
    k_xc_prep<<<x->num_inputs, 1024, 0, st>>>(d_in, x->L, x->dtype == CLB200_DTYPE_COMPLEX, (float *)x->d_mag.p,
                                              (double *)x->d_pre.p);
    k_xc_factors<<<dim3((n + 7) / 8, ns), 256, 0, st>>>((const float *)x->d_mag.p, (const double *)x->d_pre.p, x->L,
                                                        x->max_shift, (float *)x->d_fac.p);
    k_xc_findmax<<<ns, XC_GROUP, 0, st>>>((const float *)x->d_fac.p, n, std::min(n, XC_GROUP), x->max_shift, d_corr,
                                          d_lag);
    CLB_CUDA(cudaGetLastError());
    x->n_launch += 3;
    x->have_factors = true;
    return CLB200_OK;
}

// ============================================================= clComplexFilter ==
constexpr int CF_THREADS = 256;
constexpr int CF_OPT = 4;                            // outputs per thread (decimation 1)
constexpr int CF_TILE = CF_THREADS * CF_OPT;

__device__ __forceinline__ float2 cf_at(const float2 *hist, const float2 *in, long p, int km1, long n_in)
{
    // stream position p of [history (K-1 samples) | new samples]
    if (p < km1) return hist[p];
    p -= km1;
    return p < n_in ? in[p] : make_float2(0.f, 0.f);
}

// out[o] = sum_{i<K} rtaps[i] * x[skip + o*D + i], rtaps = reversed taps (td_FIR_complex_complex,
// lib/clComplexFilter_impl.cc:805-829; same sum order, i = 0..K-1).  Decimation 1: a CTA stages a
// tile of the stream and the taps in shared memory and every thread slides a register window
// over CF_OPT adjacent outputs; other decimations: one output per thread.
__global__ void __launch_bounds__(CF_THREADS)
k_cfir(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in, float2 *__restrict__ out,
       long n_out, const float2 *__restrict__ rtaps, int K, int D, int skip)
{
    extern __shared__ __align__(16) float2 cf_smem[];
    float2 *s_t = cf_smem;                 // K taps
    float2 *s_x = cf_smem + K;             // CF_TILE + K samples (decimation 1 only)
    const int km1 = K - 1;
    for (int i = threadIdx.x; i < K; i += CF_THREADS) s_t[i] = rtaps[i];
    if (D == 1) {
        const long ntile = (n_out + CF_TILE - 1) / CF_TILE;
        for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
            const long g0 = tile * CF_TILE;
            __syncthreads();
            for (int i = threadIdx.x; i < CF_TILE + km1; i += CF_THREADS) s_x[i] = cf_at(hist, in, g0 + i, km1, n_in);
            __syncthreads();
            const int o0 = threadIdx.x;                        // outputs o0 + j*CF_THREADS: conflict-free rows
            float2 acc[CF_OPT];
#pragma unroll
            for (int j = 0; j < CF_OPT; j++) acc[j] = make_float2(0.f, 0.f);
            for (int i = 0; i < K; i++) {
                const float2 a = s_t[i];
#pragma unroll
                for (int j = 0; j < CF_OPT; j++) {
                    const float2 b = s_x[o0 + j * CF_THREADS + i];
                    acc[j].x += (a.x * b.x) - (a.y * b.y);
                    acc[j].y += (a.x * b.y) + (a.y * b.x);
                }
            }
#pragma unroll
            for (int j = 0; j < CF_OPT; j++)
                if (g0 + o0 + j * CF_THREADS < n_out) out[g0 + o0 + j * CF_THREADS] = acc[j];
        }
        return;
    }
    __syncthreads();
    const long stride = (long)gridDim.x * CF_THREADS;
    for (long o = (long)blockIdx.x * CF_THREADS + threadIdx.x; o < n_out; o += stride) {
        const long g = skip + o * D;
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < K; i++) {
            const float2 a = s_t[i], b = cf_at(hist, in, g + i, km1, n_in);
            acc.x += (a.x * b.x) - (a.y * b.y);
            acc.y += (a.x * b.y) + (a.y * b.x);
        }
        out[o] = acc;
    }
}

__global__ void k_chist(const float2 *__restrict__ hist, const float2 *__restrict__ in, long n_in,
                        float2 *__restrict__ nhist, int km1)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < km1; i += gridDim.x * blockDim.x) {
        const long p = n_in + i;
        nhist[i] = p < km1 ? hist[p] : in[p - km1];
    }
}

struct CFilter : clb200_block {
    int decim = 1, skip = 0, cur = 0;
    std::vector<float2> taps, pending;
    bool updated = false;
    Buf d_hist[2], d_rtaps;
    cudaEvent_t hist_ready = nullptr;
    bool hist_pending = false;
    ~CFilter() override
    {
        DeviceGuard g(device);
        d_hist[0].release();
        d_hist[1].release();
        d_rtaps.release();
        if (hist_ready) cudaEventDestroy(hist_ready);
    }
};

int cf_configure(CFilter *f, const std::vector<float2> &taps)
{
    const int K = (int)taps.size();
    f->taps = taps;
    f->skip = 0;
    f->cur = 0;
    std::vector<float2> r(K);
    for (int i = 0; i < K; i++) r[i] = taps[K - 1 - i];
    CLB_TRY(f->d_rtaps.reserve(sizeof(float2) * K));
    CLB_CUDA(cudaMemcpy(f->d_rtaps.p, r.data(), sizeof(float2) * K, cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; i++) {
        CLB_TRY(f->d_hist[i].reserve(sizeof(float2) * std::max(1, K - 1)));
        CLB_CUDA(cudaMemset(f->d_hist[i].p, 0, sizeof(float2) * std::max(1, K - 1)));
    }
    const size_t smem = sizeof(float2) * ((size_t)2 * K + CF_TILE);
    CLB_CHECK(smem <= 200 * 1024, CLB200_EINVAL, "clComplexFilter: %d taps do not fit shared memory", K);
    // per-function limit, shared by all handles: the cap, not this filter's size
    CLB_CUDA(cudaFuncSetAttribute((const void *)k_cfir, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    f->hist_pending = false;
    return CLB200_OK;
}

int cf_apply_pending(CFilter *f)
{
    std::vector<float2> t;
    {
        std::lock_guard<std::mutex> g(f->mtx);
        if (!f->updated) return CLB200_OK;
        t = f->pending;
        f->updated = false;
    }
    return cf_configure(f, t);
}

int cf_launch(CFilter *f, const float2 *d_in, long n_in, float2 *d_out, long *n_out, cudaStream_t st)
{
    const int K = (int)f->taps.size(), D = f->decim, km1 = K - 1;
    const long nout = n_in > f->skip ? (n_in - f->skip + D - 1) / D : 0;
    if (n_out) *n_out = nout;
    if (n_in <= 0) return CLB200_OK;
    const int sms = device_sm_count(f->device);
    if (f->hist_pending) CLB_CUDA(cudaStreamWaitEvent(st, f->hist_ready, 0));
    const float2 *hist = (const float2 *)f->d_hist[f->cur].p;
    float2 *nhist = (float2 *)f->d_hist[f->cur ^ 1].p;
    if (nout > 0) {
        const size_t smem = sizeof(float2) * ((size_t)2 * K + CF_TILE);
        const long ctas = D == 1 ? (nout + CF_TILE - 1) / CF_TILE : (nout + CF_THREADS - 1) / CF_THREADS;
        k_cfir<<<grid_for(ctas, sms, 4), CF_THREADS, smem, st>>>(hist, d_in, n_in, d_out, nout,
                                                                 (const float2 *)f->d_rtaps.p, K, D, f->skip);
        CLB_CUDA(cudaGetLastError());
        f->n_launch++;
    }
    if (km1 > 0) {
        k_chist<<<(km1 + 255) / 256, 256, 0, st>>>(hist, d_in, n_in, nhist, km1);
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

// =========================================================== clQuadratureDemod ==
// quadDemod, double + fma branch (lib/clQuadratureDemod_impl.cc:126-143).  prev = the sample
// before in[0] (device-resident between calls: set_history(2), :81).
__global__ void k_quaddemod(const float2 *__restrict__ prev, const float2 *__restrict__ in, float *__restrict__ out,
                            long n, float gain, int use_gain)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 a = in[i], b = (i == 0) ? *prev : in[i - 1];
        const double a_r = a.x, a_i = a.y, b_r = b.x, b_i = -1.0 * (double)b.y;
        const double re = fma(a_r, b_r, -(a_i * b_i));
        const double im = fma(a_r, b_i, a_i * b_r);
        out[i] = use_gain ? (float)((double)gain * atan2(im, re)) : (float)atan2(im, re);
    }
}
__global__ void k_qd_keep(const float2 *__restrict__ in, long n, float2 *__restrict__ prev)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && n > 0) *prev = in[n - 1];
}

struct QuadDemod : clb200_block {
    float gain = 1.0f;
    Buf d_prev[2];
    int cur = 0;
    // consecutive launches may sit on different streams (the chunks of one work() call rotate over the
    // handle's slot streams): the kernel that reads d_prev waits for the k_qd_keep that wrote it
    cudaEvent_t prev_ready = nullptr;
    bool prev_pending = false;
    ~QuadDemod() override
    {
        DeviceGuard g(device);
        d_prev[0].release();
        d_prev[1].release();
        if (prev_ready) cudaEventDestroy(prev_ready);
    }
};

int qd_launch(QuadDemod *q, const float2 *d_in, float *d_out, long n, cudaStream_t st)
{
    if (n <= 0) return CLB200_OK;
    const int sms = device_sm_count(q->device);
    if (q->prev_pending) CLB_CUDA(cudaStreamWaitEvent(st, q->prev_ready, 0));
    k_quaddemod<<<grid_for((n + 255) / 256, sms, 8), 256, 0, st>>>((const float2 *)q->d_prev[q->cur].p, d_in, d_out, n,
                                                                   q->gain, q->gain != 1.0f);
    k_qd_keep<<<1, 32, 0, st>>>(d_in, n, (float2 *)q->d_prev[q->cur ^ 1].p);
    CLB_CUDA(cudaGetLastError());
    if (!q->prev_ready) CLB_CUDA(cudaEventCreateWithFlags(&q->prev_ready, cudaEventDisableTiming));
    CLB_CUDA(cudaEventRecord(q->prev_ready, st));
    q->prev_pending = true;
    q->cur ^= 1;
    q->n_launch += 2;
    return CLB200_OK;
}

// ============================================================== clSignalSource ==
// sig_float / sig_complex, double branch (lib/clSignalSource_impl.cc:128-211)
template <int MODE>        // 0 cos (float), 1 sin (float), 2 complex
__global__ void k_sigsource(void *__restrict__ out, long n, double phase, double phase_inc, double ampl)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double d = phase + (phase_inc * (double)i);
        if (MODE == 2) {
            double s, c;
            sincos(d, &s, &c);
            reinterpret_cast<float2 *>(out)[i] = make_float2((float)(c * ampl), (float)(s * ampl));
        } else {
            reinterpret_cast<float *>(out)[i] = (float)((MODE == 1 ? sin(d) : cos(d)) * ampl);
        }
    }
}

struct SigSource : clb200_block {
    int dtype = 0, waveform = CLB200_SIG_COS;
    double phase = 0.0, phase_inc = 0.0, ampl = 1.0;
};

int ss_launch(SigSource *s, void *d_out, long n, cudaStream_t st)
{
    if (n <= 0) return CLB200_OK;
    const int grid = grid_for((n + 255) / 256, device_sm_count(s->device), 8);
    if (s->dtype == CLB200_DTYPE_COMPLEX) k_sigsource<2><<<grid, 256, 0, st>>>(d_out, n, s->phase, s->phase_inc, s->ampl);
    else if (s->waveform == CLB200_SIG_SIN) k_sigsource<1><<<grid, 256, 0, st>>>(d_out, n, s->phase, s->phase_inc, s->ampl);
    else k_sigsource<0><<<grid, 256, 0, st>>>(d_out, n, s->phase, s->phase_inc, s->ampl);
    CLB_CUDA(cudaGetLastError());
    s->n_launch++;
    // phase bookkeeping of processOpenCL (:386-398)
    const double two_pi = 6.28318530717958647692;
    s->phase = s->phase + (s->phase_inc * (float)n);
    if (s->phase > two_pi || s->phase < -two_pi) {
        s->phase = s->phase / two_pi - (double)((int)(s->phase / two_pi));
        s->phase = s->phase * two_pi;
    }
    return CLB200_OK;
}

int need_device(int device)
{
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    return CLB200_OK;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- clXCorrelate --
int clb200_xcorrelate_create(int device, int num_inputs, int signal_length, int data_type, int max_search_index,
                             clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(num_inputs >= 2, CLB200_EINVAL, "clXCorrelate: at least 2 inputs (reference + 1), got %d", num_inputs);
    CLB_CHECK(data_type == CLB200_DTYPE_COMPLEX || data_type == CLB200_DTYPE_FLOAT, CLB200_EINVAL,
              "clXCorrelate: Unknown data type.");                                   // :710-714
    CLB_CHECK(signal_length > 0 && signal_length % 2 == 0, CLB200_EINVAL,
              "clXCorrelate: Signal length must be a multiple of 2.");               // :716-719
    CLB_CHECK(max_search_index >= 0 && max_search_index % 2 == 0, CLB200_EINVAL,
              "clXCorrelate: max shift must be a multiple of 2.");                   // :721-724
    CLB_TRY(need_device(device));
    int ms;
    if (max_search_index > 0) {
        ms = max_search_index;
    } else {
        ms = (int)(0.7f * (float)signal_length);                                      // :731-736
        if (ms % 2) ms += 1;
    }
    const int p2 = (int)pow(2.0, ceil((double)log2f((float)ms)));                     // :739-746
    CLB_CHECK(p2 >= 1 && p2 <= (1 << 24), CLB200_EINVAL, "clXCorrelate: max shift %d out of range", p2);
    DeviceGuard g(device);
    XCorr *x = new XCorr;
    x->kind = KIND_XCORR;
    x->device = device;
    x->num_inputs = num_inputs;
    x->L = signal_length;
    x->dtype = data_type;
    x->max_shift = p2;
    *out = x;
    return CLB200_OK;
}

int clb200_xcorrelate_max_shift(clb200_handle h)
{
    XCorr *x;
    if (check_kind(h, KIND_XCORR, &x) != CLB200_OK) return CLB200_EINVAL;
    return x->max_shift;
}

int clb200_xcorrelate_launch_device(clb200_handle h, const void *d_in, float *d_corr, int32_t *d_lag, void *stream)
{
    XCorr *x;
    CLB_TRY(check_kind(h, KIND_XCORR, &x));
    CLB_CHECK(d_in && d_corr && d_lag, CLB200_EINVAL, "null buffer");
    DeviceGuard g(x->device);
    return xc_launch(x, d_in, d_corr, d_lag, (cudaStream_t)stream);
}

int clb200_xcorrelate_work(clb200_handle h, const void *const *in, float *corr, int32_t *lag)
{
    XCorr *x;
    CLB_TRY(check_kind(h, KIND_XCORR, &x));
    CLB_CHECK(in && corr && lag, CLB200_EINVAL, "null buffer");
    DeviceGuard g(x->device);
    if (!x->st) CLB_CUDA(cudaStreamCreateWithFlags(&x->st, cudaStreamNonBlocking));
    const int ns = x->num_inputs - 1;
    const size_t per = x->item() * (size_t)x->L;
    CLB_TRY(x->d_in.reserve(per * x->num_inputs));
    CLB_TRY(x->d_res.reserve(8 * (size_t)ns));
    x->pin_res.host = true;
    CLB_TRY(x->pin_res.reserve(8 * (size_t)ns));
    for (int k = 0; k < x->num_inputs; k++) {
        CLB_CHECK(in[k] != nullptr, CLB200_EINVAL, "null input %d", k);
        CLB_CUDA(cudaMemcpyAsync((char *)x->d_in.p + per * k, in[k], per, cudaMemcpyHostToDevice, x->st));
        x->n_h2d += per;
    }
    float *d_corr = (float *)x->d_res.p;
    int *d_lag = (int *)((char *)x->d_res.p + 4 * (size_t)ns);
    CLB_TRY(xc_launch(x, x->d_in.p, d_corr, d_lag, x->st));
    CLB_CUDA(cudaMemcpyAsync(x->pin_res.p, x->d_res.p, 8 * (size_t)ns, cudaMemcpyDeviceToHost, x->st));
    CLB_CUDA(cudaStreamSynchronize(x->st));
    memcpy(corr, x->pin_res.p, 4 * (size_t)ns);
    memcpy(lag, (char *)x->pin_res.p + 4 * (size_t)ns, 4 * (size_t)ns);
    x->n_d2h += 8 * (size_t)ns;
    return CLB200_OK;
}

int clb200_xcorrelate_factors(clb200_handle h, int signal, float *out, int cap)
{
    XCorr *x;
    CLB_TRY(check_kind(h, KIND_XCORR, &x));
    CLB_CHECK(x->have_factors, CLB200_ESTATE, "clXCorrelate: no frame has been correlated yet");
    CLB_CHECK(signal >= 1 && signal < x->num_inputs, CLB200_EINVAL, "clXCorrelate: signal %d out of range", signal);
    const int n = 2 * x->max_shift;
    CLB_CHECK(out && cap >= n, CLB200_EINVAL, "clXCorrelate: factor buffer too small (%d < %d)", cap, n);
    DeviceGuard g(x->device);
    if (x->st) CLB_CUDA(cudaStreamSynchronize(x->st));
    CLB_CUDA(cudaMemcpy(out, (const float *)x->d_fac.p + (size_t)(signal - 1) * n, sizeof(float) * n,
                        cudaMemcpyDeviceToHost));
    return CLB200_OK;
}

// ------------------------------------------------------------- clComplexFilter --
int clb200_cfilter_create(int device, int decimation, const float *taps_c32, int ntaps, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(decimation >= 1, CLB200_EINVAL, "clComplexFilter: decimation must be >= 1, got %d", decimation);
    CLB_CHECK(taps_c32 != nullptr && ntaps >= 1, CLB200_EINVAL, "clComplexFilter: at least one tap is required");
    CLB_TRY(need_device(device));
    DeviceGuard g(device);
    CFilter *f = new CFilter;
    f->kind = KIND_CFILTER;
    f->device = device;
    f->decim = decimation;
    const float2 *t = reinterpret_cast<const float2 *>(taps_c32);
    int rc = cf_configure(f, std::vector<float2>(t, t + ntaps));
    if (rc != CLB200_OK) {
        delete f;
        return rc;
    }
    *out = f;
    return CLB200_OK;
}

int clb200_cfilter_set_taps(clb200_handle h, const float *taps_c32, int ntaps)
{
    CFilter *f;
    CLB_TRY(check_kind(h, KIND_CFILTER, &f));
    CLB_CHECK(taps_c32 != nullptr && ntaps >= 1, CLB200_EINVAL, "clComplexFilter: at least one tap is required");
    std::lock_guard<std::mutex> g(f->mtx);
    const float2 *t = reinterpret_cast<const float2 *>(taps_c32);
    f->pending.assign(t, t + ntaps);
    f->updated = true;
    return CLB200_OK;
}

int clb200_cfilter_launch_device(clb200_handle h, const void *d_in, long n_in, void *d_out, long *n_out, void *stream)
{
    CFilter *f;
    CLB_TRY(check_kind(h, KIND_CFILTER, &f));
    CLB_CHECK(n_in >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(f->device);
    CLB_TRY(cf_apply_pending(f));
    return cf_launch(f, (const float2 *)d_in, n_in, (float2 *)d_out, n_out, (cudaStream_t)stream);
}

int clb200_cfilter_work(clb200_handle h, const void *in, long n_in, void *out, long *n_out)
{
    CFilter *f;
    CLB_TRY(check_kind(h, KIND_CFILTER, &f));
    CLB_CHECK(n_in >= 0, CLB200_EINVAL, "negative item count");
    if (n_out) *n_out = 0;
    DeviceGuard g(f->device);
    CLB_TRY(cf_apply_pending(f));
    if (n_in == 0) return CLB200_OK;
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = pd.out_bytes[0] = 8;
    return run_chunked(f, pd, n_in, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *no) {
                           return cf_launch(f, (const float2 *)di[0], n, (float2 *)dout[0], no, st);
                       },
                       n_out);
}

// ----------------------------------------------------------- clQuadratureDemod --
int clb200_quaddemod_create(int device, float gain, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_TRY(need_device(device));
    DeviceGuard g(device);
    QuadDemod *q = new QuadDemod;
    q->kind = KIND_QUADDEMOD;
    q->device = device;
    q->gain = gain;
    for (int i = 0; i < 2; i++) {
        if (q->d_prev[i].reserve(sizeof(float2)) != CLB200_OK || cudaMemset(q->d_prev[i].p, 0, sizeof(float2)) != cudaSuccess) {
            delete q;
            set_error("clQuadratureDemod: device allocation failed");
            return CLB200_ENOMEM;
        }
    }
    *out = q;
    return CLB200_OK;
}

int clb200_quaddemod_launch_device(clb200_handle h, const void *d_in, void *d_out, long nitems, void *stream)
{
    QuadDemod *q;
    CLB_TRY(check_kind(h, KIND_QUADDEMOD, &q));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(q->device);
    return qd_launch(q, (const float2 *)d_in, (float *)d_out, nitems, (cudaStream_t)stream);
}

int clb200_quaddemod_work(clb200_handle h, const void *in, void *out, long nitems)
{
    QuadDemod *q;
    CLB_TRY(check_kind(h, KIND_QUADDEMOD, &q));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    if (nitems == 0) return CLB200_OK;
    DeviceGuard g(q->device);
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = 8;
    pd.out_bytes[0] = 4;
    return run_chunked(q, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           return qd_launch(q, (const float2 *)di[0], (float *)dout[0], n, st);
                       });
}

// -------------------------------------------------------------- clSignalSource --
int clb200_sigsource_create(int device, int data_type, double samp_rate, int waveform, double freq, double amplitude,
                            clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(data_type == CLB200_DTYPE_COMPLEX || data_type == CLB200_DTYPE_FLOAT, CLB200_EINVAL,
              "clSignalSource: data type must be complex or float, got %d", data_type);
    CLB_CHECK(waveform == CLB200_SIG_COS || waveform == CLB200_SIG_SIN, CLB200_EINVAL,
              "clSignalSource: waveform must be 1 (cos) or 2 (sin), got %d", waveform);
    CLB_CHECK(samp_rate > 0.0, CLB200_EINVAL, "clSignalSource: sample rate must be positive");
    CLB_TRY(need_device(device));
    SigSource *s = new SigSource;
    s->kind = KIND_SIGSOURCE;
    s->device = device;
    s->dtype = data_type;
    s->waveform = waveform;
    s->ampl = (double)(float)amplitude;                                   // the reference's amplitude is a float (:72-73)
    s->phase_inc = 6.28318530717958647692 * freq / samp_rate;             // :258
    *out = s;
    return CLB200_OK;
}

double clb200_sigsource_phase(clb200_handle h)
{
    SigSource *s;
    if (check_kind(h, KIND_SIGSOURCE, &s) != CLB200_OK) return 0.0;
    return s->phase;
}

int clb200_sigsource_launch_device(clb200_handle h, void *d_out, long nitems, void *stream)
{
    SigSource *s;
    CLB_TRY(check_kind(h, KIND_SIGSOURCE, &s));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(s->device);
    return ss_launch(s, d_out, nitems, (cudaStream_t)stream);
}

int clb200_sigsource_work(clb200_handle h, void *out, long nitems)
{
    SigSource *s;
    CLB_TRY(check_kind(h, KIND_SIGSOURCE, &s));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    if (nitems == 0) return CLB200_OK;
    DeviceGuard g(s->device);
    PortDesc pd;
    pd.nin = 0;
    pd.nout = 1;
    pd.out[0] = out;
    pd.out_bytes[0] = s->dtype == CLB200_DTYPE_COMPLEX ? 8 : 4;
    // every chunk continues the phase of the previous one: ss_launch advances it per launch exactly as
    // the reference does per work() call, so chunking changes nothing but the rollover points
    return run_chunked(s, pd, nitems, nitems,
                       [&](const void **, void **dout, long n, cudaStream_t st, long *) {
                           return ss_launch(s, dout[0], n, st);
                       });
}

} // extern "C"
