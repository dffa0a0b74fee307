// pfb.cu -- clPolyphaseChannelizer: polyphase arms + M-point inverse DFT + channel map.
//
// Reference: clPolyphaseChannelizer_impl::general_work (lib/clPolyphaseChannelizer_impl.cc:83-109)
// runs three launches per call -- filterpfb2 (:156-167), a batched clFFT BACKWARD
// plan of size num_channels (:208-225, enqueued at :100) and channel_map
// (:169-177) -- with the M-wide intermediate going through global memory twice.
// Here one kernel does all three per output time step i (T = ntaps, M = channels,
// R = inputs per iteration):
//     filt[(j + i*(M-R)) mod M] = sum_{k = j, j+M, .. < T} in[i*R - k + T-1] * taps[k]
//     fft[c]  = sum_n filt[n] e^{+2 pi i n c / M}              (unnormalised)
//     out[i*nmap + q] = fft[map[q]]
// The arm sums are computed straight into the first-pass registers of the FFT
// (each thread evaluates the arms its butterfly needs), so the only HBM traffic
// is R*8 B in + nmap*8 B out per time step (16 B/sample when critically sampled
// and fully mapped).  The arm accumulation uses fma in ascending k like the
// reference kernel (:163).
#include "common.cuh"
#include "fft_device.cuh"
#include <cmath>
#include <cstdlib>

using namespace clb200;
using namespace clb200::fftdev;

namespace {

// NT > 0: critically sampled (R == M, the commutator never rotates) with at most NT taps per arm --
// every thread keeps the taps of its arms in registers for the whole launch, which removes half of
// the load instructions of the arm loop (the kernel is LSU-bound: 73 % of the L1 wavefront peak).
template <int LOGM, int EPT, int BATCH, int MINB, int NT>
__global__ void __launch_bounds__((1 << LOGM) / EPT * BATCH, MINB)
k_pfb(const float2 *__restrict__ in, float2 *__restrict__ out, long niter,
      const float *__restrict__ taps, const float2 *__restrict__ tw, const int *__restrict__ map,
      int ntaps, int R, int nmap, int identity)
{
    using P = Plan<LOGM, EPT>;
    constexpr int M = P::N, T = P::T;
    constexpr int LINE = (P::SMEM_F2 > P::pad(M)) ? P::SMEM_F2 : P::pad(M);   // float2 per time step
    extern __shared__ __align__(16) float2 smem[];

    const int tb = (BATCH == 1) ? 0 : threadIdx.x / T;
    const int lt = (BATCH == 1) ? threadIdx.x : threadIdx.x % T;
    float2 *buf = smem + tb * LINE;
    const int rot_step = M - R;                 // commutator rotation per time step (mod M)
    float tr[NT > 0 ? NT : 1][EPT];
    if (NT > 0) {
#pragma unroll
        for (int g = 0; g < NT; g++)
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const int k = g * M + in_index<P, EPT>(lt, e);
                tr[g][e] = k < ntaps ? __ldg(taps + k) : 0.f;
            }
    }

    const long ntile = (niter + BATCH - 1) / BATCH;
    for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long i = tile * BATCH + tb;
        const bool active = i < niter;
        float2 x[EPT];
        if (active) {
            const int rot = (int)((i * (long)rot_step) & (M - 1));
            const float2 *xin = in + i * (long)R + (ntaps - 1);      // in[i*R - k + T-1]
            // tap loop outside, arm loop (unrolled) inside: EPT independent loads in flight
            // per step; each arm still accumulates in ascending k like the reference (:163)
            int j[EPT];
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                j[e] = (in_index<P, EPT>(lt, e) - rot) & (M - 1);
                x[e] = make_float2(0.f, 0.f);
            }
            if (NT > 0) {
#pragma unroll
                for (int g = 0; g < NT; g++) {
                    float2 v[EPT];
#pragma unroll
                    for (int e = 0; e < EPT; e++) {
                        const int k = g * M + j[e];
                        v[e] = k < ntaps ? __ldg(xin - k) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int e = 0; e < EPT; e++) {
                        x[e].y = fmaf(v[e].x, tr[g][e], x[e].y);
                        x[e].x = fmaf(v[e].y, tr[g][e], x[e].x);
                    }
                }
            } else
            for (int k0 = 0; k0 < ntaps; k0 += M) {
                float2 v[EPT];
                float t[EPT];
#pragma unroll
                for (int e = 0; e < EPT; e++) {
                    const int k = k0 + j[e];
                    const bool ok = k < ntaps;
                    v[e] = ok ? __ldg(xin - k) : make_float2(0.f, 0.f);
                    t[e] = ok ? __ldg(taps + k) : 0.f;
                }
#pragma unroll
                for (int e = 0; e < EPT; e++) {      // x holds (im, re): inverse via the forward core
                    x[e].y = fmaf(v[e].x, t[e], x[e].y);
                    x[e].x = fmaf(v[e].y, t[e], x[e].x);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
        }

        fft_core<P, EPT>(x, buf, lt, tw);

        if (identity) {
            if (active) {
                float2 *dst = out + i * (long)M;
                for_each_output<P, EPT>(x, lt, [&](int o, float2 a) {
                    __stcs(dst + o, make_float2(a.y, a.x));
                });
            }
        } else {
            __syncthreads();        // the last pass' reads of buf are done
            for_each_output<P, EPT>(x, lt, [&](int o, float2 a) { buf[P::pad(o)] = make_float2(a.y, a.x); });
            __syncthreads();
            if (active) {
                float2 *dst = out + i * (long)nmap;
                for (int q = lt; q < nmap; q += T) dst[q] = buf[P::pad(__ldg(map + q))];
            }
            __syncthreads();        // before the next time step overwrites buf
        }
    }
}

// Critically sampled (R == M) with NT <= 4 taps per arm: the sample an arm multiplies by tap group g at time step i is
// the one it multiplies by group 0 at step i - g.  A thread group therefore walks PFB_RUN CONSECUTIVE time steps and
// keeps the last NT - 1 steps' samples in registers: every input sample is loaded once instead of NT times (the
// kernel is LSU-bound: L1 data-pipe wavefronts 91 % busy in profiles/r2_pfb_full.txt); the next step's samples are in
// flight while a step's transform runs.  Used where it measured faster (clb200_pfb_create).  Tiles (BATCH x PFB_RUN
// steps) come from the work counter when the CTA is a single warp (common.cuh: tile_fetch), else by static striding.
constexpr int PFB_RUN = 16;
template <int LOGM, int EPT, int BATCH, int MINB, int NT>
__global__ void __launch_bounds__((1 << LOGM) / EPT * BATCH, MINB)
k_pfb_run(const float2 *__restrict__ in, float2 *__restrict__ out, long niter,
          const float *__restrict__ taps, const float2 *__restrict__ tw, const int *__restrict__ map,
          int ntaps, int nmap, int identity, unsigned long long *wq)
{
    using P = Plan<LOGM, EPT>;
    constexpr int M = P::N, T = P::T;
    constexpr int LINE = (P::SMEM_F2 > P::pad(M)) ? P::SMEM_F2 : P::pad(M);
    constexpr bool one_warp = T * BATCH == 32;
    extern __shared__ __align__(16) float2 smem[];
    const int tb = (BATCH == 1) ? 0 : threadIdx.x / T;
    const int lt = (BATCH == 1) ? threadIdx.x : threadIdx.x % T;
    float2 *buf = smem + tb * LINE;
    float tr[NT][EPT];
    bool ok[NT][EPT];
#pragma unroll
    for (int g = 0; g < NT; g++)
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int k = g * M + in_index<P, EPT>(lt, e);
            ok[g][e] = k < ntaps;
            tr[g][e] = ok[g][e] ? __ldg(taps + k) : 0.f;
        }
    // sample of arm slot e at time step i: in[i*M + ntaps-1 - j[e]]
    const float2 *const base = in + (ntaps - 1) - lt;
    const long ntile = (niter + (long)BATCH * PFB_RUN - 1) / ((long)BATCH * PFB_RUN);
    for (long tile = blockIdx.x; tile < ntile;) {
        long nxt = tile + gridDim.x;
        if (one_warp && wq != nullptr && threadIdx.x == 0) nxt = tile_fetch(wq);
        const long i0 = (tile * BATCH + tb) * PFB_RUN;
        float2 w[NT][EPT];
        // the NT - 1 steps in front of the run (history samples precede the first step)
#pragma unroll
        for (int g = 1; g < NT; g++)
#pragma unroll
            for (int e = 0; e < EPT; e++)
                w[g][e] = (ok[g][e] && i0 < niter) ? __ldg(base + (i0 - g) * (long)M - (in_index<P, EPT>(0, e)))
                                                   : make_float2(0.f, 0.f);
        // the next step's samples are in flight while this step's transform runs
        float2 nx[EPT];
#pragma unroll
        for (int e = 0; e < EPT; e++)
            nx[e] = (ok[0][e] && i0 < niter) ? __ldg(base + i0 * (long)M - (in_index<P, EPT>(0, e))) : make_float2(0.f, 0.f);
#pragma unroll 1
        for (int s = 0; s < PFB_RUN; s++) {
            const long i = i0 + s;
            const bool active = i < niter;
            float2 x[EPT];
#pragma unroll
            for (int e = 0; e < EPT; e++) w[0][e] = nx[e];
            {
                const bool more = s + 1 < PFB_RUN && i + 1 < niter;
#pragma unroll
                for (int e = 0; e < EPT; e++)
                    nx[e] = (ok[0][e] && more) ? __ldg(base + (i + 1) * (long)M - (in_index<P, EPT>(0, e))) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < NT; g++)
#pragma unroll
                for (int e = 0; e < EPT; e++) {      // ascending k like the reference (:163); x holds (im, re): inverse via the forward core
                    x[e].y = fmaf(w[g][e].x, tr[g][e], x[e].y);
                    x[e].x = fmaf(w[g][e].y, tr[g][e], x[e].x);
                }
#pragma unroll
            for (int g = NT - 1; g >= 1; g--)
#pragma unroll
                for (int e = 0; e < EPT; e++) w[g][e] = w[g - 1][e];

            fft_core<P, EPT>(x, buf, lt, tw);

            if (identity) {
                if (active) {
                    float2 *dst = out + i * (long)M;
                    for_each_output<P, EPT>(x, lt, [&](int o, float2 a) { __stcs(dst + o, make_float2(a.y, a.x)); });
                }
            } else {
                __syncthreads();        // the last pass' reads of buf are done
                for_each_output<P, EPT>(x, lt, [&](int o, float2 a) { buf[P::pad(o)] = make_float2(a.y, a.x); });
                __syncthreads();
                if (active) {
                    float2 *dst = out + i * (long)nmap;
                    for (int q = lt; q < nmap; q += T) dst[q] = buf[P::pad(__ldg(map + q))];
                }
                __syncthreads();        // before the next time step overwrites buf
            }
        }
        if (one_warp && wq != nullptr) nxt = __shfl_sync(0xffffffffu, nxt, 0);
        tile = nxt;
    }
    if (one_warp && wq != nullptr && threadIdx.x == 0) tile_finish(wq);
}

typedef void (*pfb_kernel_t)(const float2 *, float2 *, long, const float *, const float2 *,
                             const int *, int, int, int, int);
typedef void (*pfb_run_kernel_t)(const float2 *, float2 *, long, const float *, const float2 *,
                                 const int *, int, int, int, unsigned long long *);
struct PfbVariant {
    int logm, batch, threads, smem_bytes;
    void (*fill_tw)(std::vector<float2> &);
    pfb_kernel_t kernel;          // any R, any tap count
    pfb_kernel_t kernel_crit[2];  // R == M and <= 2 / <= 4 taps per arm: taps in registers
    pfb_run_kernel_t kernel_run[2];   // the same with runs of consecutive time steps per thread group (samples loaded once)
};

template <int LOGM, int EPT>
void fill_tw_p(std::vector<float2> &tw)
{
    using P = Plan<LOGM, EPT>;
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

template <int LOGM, int EPT, int BATCH, int MINB>
PfbVariant make_pfb()
{
    using P = Plan<LOGM, EPT>;
    constexpr int LINE = (P::SMEM_F2 > P::pad(P::N)) ? P::SMEM_F2 : P::pad(P::N);
    return PfbVariant{LOGM, BATCH, P::T * BATCH, LINE * BATCH * (int)sizeof(float2),
                      &fill_tw_p<LOGM, EPT>, &k_pfb<LOGM, EPT, BATCH, MINB, 0>,
                      {&k_pfb<LOGM, EPT, BATCH, MINB, 2>, &k_pfb<LOGM, EPT, BATCH, MINB, 4>},
                      {&k_pfb_run<LOGM, EPT, BATCH, (MINB > 16 ? 16 : MINB), 2>, &k_pfb_run<LOGM, EPT, BATCH, (MINB > 16 ? 16 : MINB), 4>}};
}

const PfbVariant *pick_pfb(int logm)
{
    static const PfbVariant tab[] = {
        // up to 128 channels: one-warp CTAs (barriers are free, many CTAs hide the load
        // latency) -- measured best for 64 channels: 4.36 TB/s vs 3.67 with 256 threads
        make_pfb<1, 2, 32, 24>(),  make_pfb<2, 4, 32, 24>(),  make_pfb<3, 8, 32, 16>(),
        make_pfb<4, 4, 8, 24>(),   make_pfb<5, 8, 8, 24>(),   make_pfb<6, 8, 4, 24>(),
        make_pfb<7, 8, 2, 24>(),   make_pfb<8, 16, 16, 2>(),  make_pfb<9, 8, 4, 4>(),
        make_pfb<10, 16, 4, 2>(),  make_pfb<11, 16, 2, 2>(),  make_pfb<12, 16, 1, 2>(),
    };
    if (logm < 1 || logm > 12) return nullptr;
    return &tab[logm - 1];
}

struct Pfb : clb200_block {
    int ntaps = 0, M = 0, R = 0, nmap = 0, buf_items = 0, identity = 0, resident = 1;
    const PfbVariant *var = nullptr;
    pfb_kernel_t kernel = nullptr;
    pfb_run_kernel_t kernel_run = nullptr;   // set: the run kernel replaces `kernel`
    Buf d_taps, d_tw, d_map;
    ~Pfb() override
    {
        DeviceGuard g(device);
        d_taps.release();
        d_tw.release();
        d_map.release();
    }
};

int pfb_launch(Pfb *p, const void *d_in, void *d_out, long niter, cudaStream_t st)
{
    if (niter <= 0) return CLB200_OK;
    const PfbVariant *v = p->var;
    if (p->kernel_run) {
        const long per = (long)v->batch * PFB_RUN;
        const long ntile = (niter + per - 1) / per;
        const int grid = grid_for(ntile, device_sm_count(p->device), p->resident);
        p->kernel_run<<<grid, v->threads, v->smem_bytes, st>>>(
            (const float2 *)d_in, (float2 *)d_out, niter, (const float *)p->d_taps.p, (const float2 *)p->d_tw.p,
            (const int *)p->d_map.p, p->ntaps, p->nmap, p->identity,
            (v->threads == 32 && ntile > grid) ? p->work_counter(st) : nullptr);
        CLB_CUDA(cudaGetLastError());
        p->n_launch++;
        return CLB200_OK;
    }
    long ntile = (niter + v->batch - 1) / v->batch;
    int grid = grid_for(ntile, device_sm_count(p->device), p->resident);
    p->kernel<<<grid, v->threads, v->smem_bytes, st>>>(
        (const float2 *)d_in, (float2 *)d_out, niter, (const float *)p->d_taps.p,
        (const float2 *)p->d_tw.p, (const int *)p->d_map.p, p->ntaps, p->R, p->nmap, p->identity);
    CLB_CUDA(cudaGetLastError());
    p->n_launch++;
    return CLB200_OK;
}

} // namespace

extern "C" {

int clb200_pfb_create(int device, const float *taps, int ntaps, int buf_items, int num_channels,
                      int ninputs_per_iter, const int *ch_map, int nmap, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(taps != nullptr && ntaps >= 1, CLB200_EINVAL, "clPolyphaseChannelizer: taps required");
    const int M = num_channels, R = ninputs_per_iter;
    CLB_CHECK(M >= 2 && (M & (M - 1)) == 0 && M <= 4096, CLB200_EINVAL,
              "clPolyphaseChannelizer: num_channels %d is not a power of two in 2..4096", M);
    CLB_CHECK(R >= 1 && R <= M, CLB200_EINVAL,
              "clPolyphaseChannelizer: ninputs_per_iter %d must be in 1..num_channels", R);
    // lib/clPolyphaseChannelizer_impl.cc:59-62 throws std::invalid_argument with this text
    CLB_CHECK(buf_items >= M && buf_items % M == 0, CLB200_EINVAL,
              "buf_items must be a multiple of num_channels (%d vs %d)", buf_items, M);
    CLB_CHECK(buf_items % R == 0, CLB200_EINVAL,
              "clPolyphaseChannelizer: buf_items %d must be a positive multiple of ninputs_per_iter %d",
              buf_items, R);
    CLB_CHECK(ch_map != nullptr && nmap >= 1, CLB200_EINVAL, "clPolyphaseChannelizer: empty channel map");
    for (int q = 0; q < nmap; q++)
        CLB_CHECK(ch_map[q] >= 0 && ch_map[q] < M, CLB200_EINVAL,
                  "clPolyphaseChannelizer: ch_map[%d]=%d outside 0..%d", q, ch_map[q], M - 1);
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    DeviceGuard g(device);
    Pfb *p = new Pfb;
    p->kind = KIND_PFB;
    p->device = device;
    p->init_work_counters();
    p->ntaps = ntaps;
    p->M = M;
    p->R = R;
    p->nmap = nmap;
    p->buf_items = buf_items;
    p->identity = (nmap == M);
    for (int q = 0; q < nmap && p->identity; q++) p->identity = (ch_map[q] == q);
    p->var = pick_pfb(ilog2(M));
    auto fail = [&](int rc) {
        delete p;
        return rc;
    };
    std::vector<float2> tw;
    p->var->fill_tw(tw);
    if (p->d_taps.reserve(sizeof(float) * ntaps) || p->d_tw.reserve(sizeof(float2) * tw.size()) ||
        p->d_map.reserve(sizeof(int) * nmap))
        return fail(CLB200_ENOMEM);
    if (cudaMemcpy(p->d_taps.p, taps, sizeof(float) * ntaps, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_tw.p, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_map.p, ch_map, sizeof(int) * nmap, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("clPolyphaseChannelizer: table upload failed");
        return fail(CLB200_ECUDA);
    }
    p->kernel = p->var->kernel;
    {
        const char *gen = getenv("CLB200_PFB_GENERAL");           // A/B: force the general kernel
        const int per_arm = (ntaps + M - 1) / M;
        if (R == M && per_arm <= 4 && !(gen && atoi(gen))) {
            p->kernel = p->var->kernel_crit[per_arm <= 2 ? 0 : 1];
            // runs of consecutive time steps: measured (tools/pfb_ab.py) +12 % at 64 channels x 4 taps per arm and +47 % at
            // 16 x 4, but -7 % at 64 x 2 (the BASELINE shape: the second read of a sample is an L1 hit there and the run's
            // serial steps cost more than it saves) and -14 ... -18 % with multi-warp CTAs: one-warp CTAs with 3-4 taps per arm
            const char *run = getenv("CLB200_PFB_RUN");           // A/B: 0 = never, 1 = whenever the kernel exists
            const bool want = run ? atoi(run) != 0 : (p->var->threads == 32 && per_arm > 2);
            if (want) p->kernel_run = p->var->kernel_run[per_arm <= 2 ? 0 : 1];
        }
    }
    const void *kfn = p->kernel_run ? (const void *)p->kernel_run : (const void *)p->kernel;
    cudaError_t e = cudaFuncSetAttribute(kfn,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         p->var->smem_bytes);
    int occ = 0;
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, p->var->threads, p->var->smem_bytes);
    if (e != cudaSuccess || occ < 1) {
        set_error("clPolyphaseChannelizer: kernel does not fit an SM (%s)", cudaGetErrorString(e));
        return fail(CLB200_ECUDA);
    }
    p->resident = occ;
    p->set_info("clPolyphaseChannelizer %d channels, %d taps, %d inputs/iteration, %d mapped outputs: %d-pt inverse FFT per time step, "
                "%d threads/CTA, %d B shared memory, %d CTAs/SM%s",
                num_channels, ntaps, ninputs_per_iter, nmap, num_channels, p->var->threads, p->var->smem_bytes, p->resident,
                p->identity ? ", identity map (direct stores)" : ", channel map through shared memory");
    *out = p;
    return CLB200_OK;
}

int clb200_pfb_launch_device(clb200_handle h, const void *d_in, void *d_out, long niter, void *stream)
{
    Pfb *p;
    CLB_TRY(check_kind(h, KIND_PFB, &p));
    CLB_CHECK(niter >= 0, CLB200_EINVAL, "negative iteration count");
    DeviceGuard g(p->device);
    return pfb_launch(p, d_in, d_out, niter, (cudaStream_t)stream);
}

int clb200_pfb_work(clb200_handle h, const void *in, void *out, long niter)
{
    Pfb *p;
    CLB_TRY(check_kind(h, KIND_PFB, &p));
    CLB_CHECK(niter >= 0, CLB200_EINVAL, "negative iteration count");
    if (niter == 0) return CLB200_OK;
    DeviceGuard g(p->device);
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = (size_t)p->R * 8;                 // R new samples per time step ...
    pd.in_extra[0] = (long)(p->ntaps - p->R) * 8;        // ... + the history overlap
    pd.out_bytes[0] = (size_t)p->nmap * 8;
    return run_chunked(p, pd, niter, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           return pfb_launch(p, di[0], dout[0], n, st);
                       });
}

} // extern "C"
