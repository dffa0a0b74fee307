// elementwise.cu -- clMathConst, clMathOp and the secondary element-wise blocks.
//
// Reference semantics (kernel strings):
//   opconst_complex / op_complex(conj) / opconst_float / opconst_int   lib/clMathConst_impl.cc:121-222
//   op_float / op_int / op_complex (mul, add, sub, mul-conj)            lib/clMathOp_impl.cc:104-238
//   op_log10 lib/clLog_impl.cc:139-148, op_snr lib/clSNR_impl.cc:105-113,
//   complextomag / complextoarg / complextomagphase / magphasetocomplex (see include/clenabled_b200.h)
//
// All of these are pure HBM streams (16 B/sample for the c32 1->1 ops): the
// kernels are grid-stride loops over 128-bit vectors, 4 independent loads in
// flight per thread, streaming cache hints, grid = SMs x resident CTAs.
// Arithmetic that must be bit-identical to the reference CPU loop uses the
// non-contracting intrinsics (__fmul_rn/__fadd_rn/__fsub_rn): the reference
// CPU path (clMathOp_impl.cc:346-349) rounds every product and sum separately.
#include "common.cuh"

using namespace clb200;

namespace {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;
constexpr int EW_CTAS_PER_SM = 8;

struct OpMul  { __device__ static float f(float a, float k) { return __fmul_rn(a, k); }
                __device__ static int   i(int a, int k) { return (int)((unsigned)a * (unsigned)k); } };
struct OpAdd  { __device__ static float f(float a, float k) { return __fadd_rn(a, k); }
                __device__ static int   i(int a, int k) { return (int)((unsigned)a + (unsigned)k); } };
struct OpSub  { __device__ static float f(float a, float k) { return __fsub_rn(a, k); }
                __device__ static int   i(int a, int k) { return (int)((unsigned)a - (unsigned)k); } };

__device__ __forceinline__ float4 ld4(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st4(float4 *p, float4 v) { __stcs(p, v); }

// Tile loop shared by the element-wise kernels: `nv` vector items in tiles of EW_THREADS x U, handed out by the work
// counter (common.cuh: tile_fetch; static striding when wq is null).  full(i) handles the U items i + u * EW_THREADS
// of a whole tile (all loads first), one(i) a single item of the ragged end (first CTA).
template <int U, class Full, class One>
__device__ __forceinline__ void ew_tile_loop(long nv, unsigned long long *wq, Full &&full, One &&one)
{
    constexpr long TILE = (long)EW_THREADS * U;
    __shared__ long s_next[2];                 // double-buffered: a slow reader of tile t never sees the index written for t + 1
    const long ntile = nv / TILE;
    int it = 0;
    for (long tile = blockIdx.x; tile < ntile; it ^= 1) {
        long nxt = tile + gridDim.x;
        if (wq != nullptr && threadIdx.x == 0) nxt = tile_fetch(wq);
        full(tile * TILE + threadIdx.x);
        if (wq != nullptr) {
            if (threadIdx.x == 0) s_next[it] = nxt;
            __syncthreads();
            nxt = s_next[it];
        }
        tile = nxt;
    }
    if (wq != nullptr && threadIdx.x == 0) tile_finish(wq);
    if (blockIdx.x == 0)
        for (long i = ntile * TILE + threadIdx.x; i < nv; i += EW_THREADS) one(i);
}

// ---- 1 in -> 1 out over float4 vectors ------------------------------------
// Tiles of EW_THREADS x EW_UNROLL vectors (16 KiB) handed out by the work counter (common.cuh: tile_fetch; static
// striding without one): the resident CTAs of an SM then all stay busy until the launch ends.
template <class F>
__global__ void __launch_bounds__(EW_THREADS)
k_map1(const float4 *__restrict__ in, float4 *__restrict__ out, long nvec, long nscalar, F f, unsigned long long *wq)
{
    constexpr long TILE = (long)EW_THREADS * EW_UNROLL;
    __shared__ long s_next[2];
    const long ntile = nvec / TILE;
    int it = 0;
    for (long tile = blockIdx.x; tile < ntile; it ^= 1) {
        long nxt = tile + gridDim.x;
        if (wq != nullptr && threadIdx.x == 0) nxt = tile_fetch(wq);
        const long i = tile * TILE + threadIdx.x;
        float4 v[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; u++) v[u] = ld4(in + i + u * EW_THREADS);
#pragma unroll
        for (int u = 0; u < EW_UNROLL; u++) st4(out + i + u * EW_THREADS, f.vec(v[u]));
        if (wq != nullptr) {
            if (threadIdx.x == 0) s_next[it] = nxt;
            __syncthreads();
            nxt = s_next[it];
        }
        tile = nxt;
    }
    if (wq != nullptr && threadIdx.x == 0) tile_finish(wq);
    // the last partial tile, then the scalar tail (nscalar % 4 floats): first CTA only
    if (blockIdx.x == 0) {
        for (long i = ntile * TILE + threadIdx.x; i < nvec; i += EW_THREADS) st4(out + i, f.vec(ld4(in + i)));
        long t = nvec * 4 + threadIdx.x;
        if (t < nscalar) {
            const float *si = reinterpret_cast<const float *>(in);
            float *so = reinterpret_cast<float *>(out);
            so[t] = f.scalar(si[t], (int)(t & 1));
        }
    }
}

// Scheduler-sized launches (the zero-copy path of run_chunked: the pointers are pinned HOST memory): one vector per
// thread, 64-thread CTAs -- an SM keeps only a few reads of system memory in flight (one CTA needs 110 us for 64 KiB,
// profiles/r2_mailbox_probe.txt), so the call's 4096 vectors go to 64 SMs instead of 4.
template <class F>
__global__ void __launch_bounds__(64)
k_map1_small(const float4 *__restrict__ in, float4 *__restrict__ out, long nvec, long nscalar, F f)
{
    const long i = (long)blockIdx.x * 64 + threadIdx.x;
    if (i < nvec) st4(out + i, f.vec(ld4(in + i)));
    if (blockIdx.x == 0) {
        long t = nvec * 4 + threadIdx.x;
        if (t < nscalar) {
            const float *si = reinterpret_cast<const float *>(in);
            float *so = reinterpret_cast<float *>(out);
            so[t] = f.scalar(si[t], (int)(t & 1));
        }
    }
}

// unaligned fallback: plain scalar grid-stride
template <class F>
__global__ void __launch_bounds__(EW_THREADS)
k_map1_scalar(const float *__restrict__ in, float *__restrict__ out, long n, F f)
{
    long stride = (long)gridDim.x * EW_THREADS;
    for (long i = (long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += stride)
        out[i] = f.scalar(in[i], (int)(i & 1));
}

template <class OP>
struct ConstF {     // float / complex: k applied to every float (clMathConst_impl.cc:169-201)
    float k;
    __device__ float4 vec(float4 v) const
    {
        return make_float4(OP::f(v.x, k), OP::f(v.y, k), OP::f(v.z, k), OP::f(v.w, k));
    }
    __device__ float scalar(float a, int) const { return OP::f(a, k); }
};
template <class OP>
struct ConstI {     // opconst_int (:144-167): integer arithmetic on the bit pattern
    int k;
    __device__ float4 vec(float4 v) const
    {
        return make_float4(__int_as_float(OP::i(__float_as_int(v.x), k)),
                           __int_as_float(OP::i(__float_as_int(v.y), k)),
                           __int_as_float(OP::i(__float_as_int(v.z), k)),
                           __int_as_float(OP::i(__float_as_int(v.w), k)));
    }
    __device__ float scalar(float a, int) const { return __int_as_float(OP::i(__float_as_int(a), k)); }
};
struct Conj {       // op_complex conjugate (:203-218): imag * -1.0
    __device__ float4 vec(float4 v) const { return make_float4(v.x, -v.y, v.z, -v.w); }
    __device__ float scalar(float a, int odd) const { return odd ? -a : a; }
};
struct Copy {
    __device__ float4 vec(float4 v) const { return v; }
    __device__ float scalar(float a, int) const { return a; }
};

template <class F>
int launch_map1(const void *in, void *out, long nfloats, F f, int sms, cudaStream_t st, clb200_block *blk = nullptr)
{
    if (nfloats <= 0) return CLB200_OK;
    bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    if (aligned) {
        long nvec = nfloats / 4;
        if (nvec > 0 && nvec <= 64L * 1024) {                    // <= 1 MiB: spread over as many SMs as there are
            k_map1_small<F><<<(unsigned)((nvec + 63) / 64), 64, 0, st>>>((const float4 *)in, (float4 *)out, nvec, nfloats, f);
            CLB_CUDA(cudaGetLastError());
            return CLB200_OK;
        }
        long ctas = (nvec + EW_THREADS * EW_UNROLL - 1) / (EW_THREADS * EW_UNROLL);
        int grid = grid_for(ctas, sms, EW_CTAS_PER_SM);
        k_map1<F><<<grid, EW_THREADS, 0, st>>>((const float4 *)in, (float4 *)out, nvec, nfloats, f,
                                               (blk && ctas > 2L * grid) ? blk->work_counter(st) : nullptr);
    } else {
        long ctas = (nfloats + EW_THREADS - 1) / EW_THREADS;
        int grid = grid_for(ctas, sms, EW_CTAS_PER_SM);
        k_map1_scalar<F><<<grid, EW_THREADS, 0, st>>>((const float *)in, (float *)out, nfloats, f);
    }
    CLB_CUDA(cudaGetLastError());
    return CLB200_OK;
}

// ---- 2 in -> 1 out ----------------------------------------------------------
struct Bin {
    int op;
    bool cplx, is_int;
    __device__ float2 cmul(float2 a, float2 b) const
    {   // (ar*br - ai*bi, ar*bi + ai*br), each op rounded (clMathOp_impl.cc:194-201,346-349)
        return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                           __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
    }
    __device__ float s(float a, float b) const
    {
        if (is_int) {
            unsigned x = (unsigned)__float_as_int(a), y = (unsigned)__float_as_int(b);
            unsigned r = op == CLB200_OP_MULTIPLY ? x * y : (op == CLB200_OP_ADD ? x + y : x - y);
            return __int_as_float((int)r);
        }
        return op == CLB200_OP_MULTIPLY ? __fmul_rn(a, b)
                                        : (op == CLB200_OP_ADD ? __fadd_rn(a, b) : __fsub_rn(a, b));
    }
    __device__ float4 vec(float4 a, float4 b) const
    {
        if (cplx && (op == CLB200_OP_MULTIPLY || op == CLB200_OP_MULTIPLY_CONJ)) {
            float sg = (op == CLB200_OP_MULTIPLY_CONJ) ? -1.0f : 1.0f;   // b_i = -1.0*b.imag (:228)
            float2 r0 = cmul(make_float2(a.x, a.y), make_float2(b.x, sg * b.y));
            float2 r1 = cmul(make_float2(a.z, a.w), make_float2(b.z, sg * b.w));
            return make_float4(r0.x, r0.y, r1.x, r1.y);
        }
        return make_float4(s(a.x, b.x), s(a.y, b.y), s(a.z, b.z), s(a.w, b.w));
    }
};

__global__ void __launch_bounds__(EW_THREADS)
k_map2(const float4 *__restrict__ a, const float4 *__restrict__ b, float4 *__restrict__ c,
       long nvec, Bin f, unsigned long long *wq)
{
    constexpr int U = 4;                                  // four vectors per operand in flight per thread
    ew_tile_loop<U>(nvec, wq,
        [&](long i) {
            float4 va[U], vb[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                va[u] = ld4(a + i + u * EW_THREADS);
                vb[u] = ld4(b + i + u * EW_THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; u++) st4(c + i + u * EW_THREADS, f.vec(va[u], vb[u]));
        },
        [&](long i) { st4(c + i, f.vec(ld4(a + i), ld4(b + i))); });
}

// element granularity fallback / tail: `pairs` = process float2 items (complex) else floats
__global__ void __launch_bounds__(EW_THREADS)
k_map2_scalar(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ c,
              long first, long n, Bin f)
{
    long stride = (long)gridDim.x * EW_THREADS;
    if (f.cplx) {
        for (long i = first / 2 + (long)blockIdx.x * EW_THREADS + threadIdx.x; i < n / 2; i += stride) {
            float4 r = f.vec(make_float4(a[2 * i], a[2 * i + 1], 0.f, 0.f),
                             make_float4(b[2 * i], b[2 * i + 1], 0.f, 0.f));
            c[2 * i] = r.x;
            c[2 * i + 1] = r.y;
        }
    } else {
        for (long i = first + (long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += stride)
            c[i] = f.s(a[i], b[i]);
    }
}

int launch_map2(const void *a, const void *b, void *c, long nfloats, Bin f, int sms, cudaStream_t st, clb200_block *blk = nullptr)
{
    if (nfloats <= 0) return CLB200_OK;
    bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
    long nvec = aligned ? nfloats / 4 : 0;
    if (nvec > 0) {
        long ctas = (nvec + EW_THREADS * 4 - 1) / (EW_THREADS * 4);
        const int grid = grid_for(ctas, sms, EW_CTAS_PER_SM);
        k_map2<<<grid, EW_THREADS, 0, st>>>((const float4 *)a, (const float4 *)b, (float4 *)c, nvec, f,
                                            (blk && ctas > 2L * grid) ? blk->work_counter(st) : nullptr);
    }
    if (nvec * 4 < nfloats) {
        long rest = nfloats - nvec * 4;
        long ctas = (rest + EW_THREADS - 1) / EW_THREADS;
        k_map2_scalar<<<grid_for(ctas, sms, EW_CTAS_PER_SM), EW_THREADS, 0, st>>>(
            (const float *)a, (const float *)b, (float *)c, nvec * 4, nfloats, f);
    }
    CLB_CUDA(cudaGetLastError());
    return CLB200_OK;
}

// ---- secondary kernels (item granularity; transcendental -> not bit-exact) ----
__global__ void __launch_bounds__(EW_THREADS)
k_log10(const float *__restrict__ a, float *__restrict__ c, long n, float factor, float k, unsigned long long *wq)
{   // c = (n/log2(10)) * log2(a) + k    (clLog_impl.cc:139-148)
    const bool vec = (((uintptr_t)a | (uintptr_t)c) & 15) == 0;
    const long nv = vec ? n / 4 : 0;                       // 16 B per access, four in flight per thread
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    float4 *c4 = reinterpret_cast<float4 *>(c);
    auto f4 = [&](float4 v) {
        return make_float4(fmaf(factor, log2f(v.x), k), fmaf(factor, log2f(v.y), k), fmaf(factor, log2f(v.z), k),
                           fmaf(factor, log2f(v.w), k));
    };
    constexpr int U = 4;
    ew_tile_loop<U>(nv, wq,
        [&](long i) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld4(a4 + i + u * EW_THREADS);
#pragma unroll
            for (int u = 0; u < U; u++) st4(c4 + i + u * EW_THREADS, f4(v[u]));
        },
        [&](long i) { st4(c4 + i, f4(ld4(a4 + i))); });
    const long stride = (long)gridDim.x * EW_THREADS;
    for (long j = nv * 4 + (long)blockIdx.x * EW_THREADS + threadIdx.x; j < n; j += stride)
        c[j] = fmaf(factor, log2f(a[j]), k);
}

__global__ void __launch_bounds__(EW_THREADS)
k_snr(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ c, long n,
      float nv, float kv)
{   // c = fabs(n*log10(a/b)+k)   (clSNR_impl.cc:105-113)
    long stride = (long)gridDim.x * EW_THREADS;
    for (long i = (long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += stride)
        c[i] = fabsf(fmaf(nv, log10f(a[i] / b[i]), kv));
}

template <bool MAG, bool ARG>
__global__ void __launch_bounds__(EW_THREADS)
k_c2mp(const float2 *__restrict__ a, float *__restrict__ mag, float *__restrict__ ph, long n, unsigned long long *wq)
{   // complextomag / complextoarg / complextomagphase
    // two samples per access (16 B in, 8 B out per output stream) when the pointers allow it
    const bool vec = (((uintptr_t)a & 15) | ((MAG ? (uintptr_t)mag : 0) & 7) | ((ARG ? (uintptr_t)ph : 0) & 7)) == 0;
    const long nv = vec ? n / 2 : 0;
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    auto put = [&](long i, float4 v) {
        if (MAG) __stcs(reinterpret_cast<float2 *>(mag) + i, make_float2(sqrtf(v.y * v.y + v.x * v.x), sqrtf(v.w * v.w + v.z * v.z)));
        if (ARG) __stcs(reinterpret_cast<float2 *>(ph) + i, make_float2(atan2f(v.y, v.x), atan2f(v.w, v.z)));
    };
    constexpr int U = 4;
    ew_tile_loop<U>(nv, wq,
        [&](long i) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld4(a4 + i + u * EW_THREADS);
#pragma unroll
            for (int u = 0; u < U; u++) put(i + u * EW_THREADS, v[u]);
        },
        [&](long i) { put(i, ld4(a4 + i)); });
    const long stride = (long)gridDim.x * EW_THREADS;
    for (long j = nv * 2 + (long)blockIdx.x * EW_THREADS + threadIdx.x; j < n; j += stride) {
        float2 v = __ldcs(a + j);
        if (MAG) mag[j] = sqrtf(v.y * v.y + v.x * v.x);
        if (ARG) ph[j] = atan2f(v.y, v.x);
    }
}

__global__ void __launch_bounds__(EW_THREADS)
k_mp2c(const float *__restrict__ mag, const float *__restrict__ ph, float2 *__restrict__ c, long n)
{   // magphasetocomplex (clMagPhaseToComplex_impl.cc:169-192)
    long stride = (long)gridDim.x * EW_THREADS;
    for (long i = (long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += stride) {
        float s, co;
        sincosf(ph[i], &s, &co);
        float m = mag[i];
        c[i] = make_float2(m * co, m * s);
    }
}

inline int item_grid(long n, int sms)
{
    return grid_for((n + EW_THREADS - 1) / EW_THREADS, sms, EW_CTAS_PER_SM);
}
// grid and work counter of the tiled kernels (tiles of EW_THREADS x 4 vector items)
inline int tile_grid(long nv, int sms)
{
    return grid_for((nv + EW_THREADS * 4 - 1) / (EW_THREADS * 4), sms, EW_CTAS_PER_SM);
}
inline unsigned long long *tile_wq(clb200_block *b, long nv, int sms, cudaStream_t st)
{
    return nv / (EW_THREADS * 4) > 2L * sms * EW_CTAS_PER_SM ? b->work_counter(st) : nullptr;
}

// ---------------------------------------------------------------- handles ----
struct MathConst : clb200_block {
    int dtype, op;
    float k;
};
struct MathOp : clb200_block {
    int dtype, op;
};
struct Unary : clb200_block {
    int ukind;
    float nv, kv;
};

int floats_per_item(int dtype) { return dtype == CLB200_DTYPE_COMPLEX ? 2 : 1; }

int mathconst_launch(MathConst *m, const void *in, void *out, long nitems, cudaStream_t st)
{
    long nf = nitems * floats_per_item(m->dtype);
    int sms = device_sm_count(m->device);
    float k;
    {
        std::lock_guard<std::mutex> g(m->mtx);
        k = m->k;
    }
    m->n_launch++;
    bool is_int = m->dtype == CLB200_DTYPE_INT;
    switch (m->op) {
    case CLB200_OP_MULTIPLY:
        return is_int ? launch_map1(in, out, nf, ConstI<OpMul>{(int)k}, sms, st, m)
                      : launch_map1(in, out, nf, ConstF<OpMul>{k}, sms, st, m);
    case CLB200_OP_ADD:
        return is_int ? launch_map1(in, out, nf, ConstI<OpAdd>{(int)k}, sms, st, m)
                      : launch_map1(in, out, nf, ConstF<OpAdd>{k}, sms, st, m);
    case CLB200_OP_SUBTRACT:
        return is_int ? launch_map1(in, out, nf, ConstI<OpSub>{(int)k}, sms, st, m)
                      : launch_map1(in, out, nf, ConstF<OpSub>{k}, sms, st, m);
    case CLB200_OP_COMPLEX_CONJ:
        return launch_map1(in, out, nf, Conj{}, sms, st, m);
    case CLB200_OP_EMPTY_W_COPY:
        // the reference's copy case falls through into the multiply (missing break,
        // clMathConst_impl.cc:187-193): not reproduced, this is a copy.
        return launch_map1(in, out, nf, Copy{}, sms, st, m);
    default:   // MATHOP_EMPTY: the kernel body is empty, output untouched
        m->n_launch--;
        return CLB200_OK;
    }
}

} // namespace

extern "C" {

int clb200_mathconst_create(int dtype, int device, float k, int op, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(dtype == CLB200_DTYPE_COMPLEX || dtype == CLB200_DTYPE_FLOAT || dtype == CLB200_DTYPE_INT,
              CLB200_EINVAL, "clMathConst: unsupported data type %d", dtype);
    bool known = op == CLB200_OP_MULTIPLY || op == CLB200_OP_ADD || op == CLB200_OP_SUBTRACT ||
                 op == CLB200_OP_COMPLEX_CONJ || op == CLB200_OP_EMPTY || op == CLB200_OP_EMPTY_W_COPY;
    CLB_CHECK(known, CLB200_EINVAL, "clMathConst: unknown operator %d", op);
    CLB_CHECK(op != CLB200_OP_COMPLEX_CONJ || dtype == CLB200_DTYPE_COMPLEX, CLB200_EINVAL,
              "clMathConst: conjugate needs complex data");
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    MathConst *m = new MathConst;
    m->kind = KIND_MATHCONST;
    m->device = device;
    m->init_work_counters();
    m->dtype = dtype;
    m->op = op;
    m->k = k;
    *out = m;
    return CLB200_OK;
}

int clb200_mathconst_set_k(clb200_handle h, float k)
{
    MathConst *m;
    CLB_TRY(check_kind(h, KIND_MATHCONST, &m));
    std::lock_guard<std::mutex> g(m->mtx);
    m->k = k;
    return CLB200_OK;
}

float clb200_mathconst_k(clb200_handle h)
{
    MathConst *m;
    if (check_kind(h, KIND_MATHCONST, &m) != CLB200_OK) return 0.0f;
    std::lock_guard<std::mutex> g(m->mtx);
    return m->k;
}

int clb200_mathconst_launch_device(clb200_handle h, const void *d_in, void *d_out, long nitems,
                                   void *stream)
{
    MathConst *m;
    CLB_TRY(check_kind(h, KIND_MATHCONST, &m));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(m->device);
    return mathconst_launch(m, d_in, d_out, nitems, (cudaStream_t)stream);
}

int clb200_mathconst_work(clb200_handle h, const void *in, void *out, long nitems)
{
    MathConst *m;
    CLB_TRY(check_kind(h, KIND_MATHCONST, &m));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    if (nitems == 0 || m->op == CLB200_OP_EMPTY) return CLB200_OK;
    DeviceGuard g(m->device);
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = pd.out_bytes[0] = 4 * floats_per_item(m->dtype);
    return run_chunked(m, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           return mathconst_launch(m, di[0], dout[0], n, st);
                       });
}

int clb200_mathop_create(int dtype, int device, int op, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(dtype == CLB200_DTYPE_COMPLEX || dtype == CLB200_DTYPE_FLOAT || dtype == CLB200_DTYPE_INT,
              CLB200_EINVAL, "clMathOp: unsupported data type %d", dtype);
    bool known = op == CLB200_OP_MULTIPLY || op == CLB200_OP_ADD || op == CLB200_OP_SUBTRACT ||
                 op == CLB200_OP_MULTIPLY_CONJ;
    CLB_CHECK(known, CLB200_EINVAL, "clMathOp: unknown operator %d", op);
    CLB_CHECK(op != CLB200_OP_MULTIPLY_CONJ || dtype == CLB200_DTYPE_COMPLEX, CLB200_EINVAL,
              "clMathOp: multiply-conjugate needs complex data");
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    MathOp *m = new MathOp;
    m->kind = KIND_MATHOP;
    m->device = device;
    m->init_work_counters();
    m->dtype = dtype;
    m->op = op;
    *out = m;
    return CLB200_OK;
}

int clb200_mathop_launch_device(clb200_handle h, const void *d_a, const void *d_b, void *d_c,
                                long nitems, void *stream)
{
    MathOp *m;
    CLB_TRY(check_kind(h, KIND_MATHOP, &m));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    DeviceGuard g(m->device);
    Bin f{m->op, m->dtype == CLB200_DTYPE_COMPLEX, m->dtype == CLB200_DTYPE_INT};
    m->n_launch++;
    return launch_map2(d_a, d_b, d_c, nitems * floats_per_item(m->dtype), f,
                       device_sm_count(m->device), (cudaStream_t)stream, m);
}

int clb200_mathop_work(clb200_handle h, const void *a, const void *b, void *c, long nitems)
{
    MathOp *m;
    CLB_TRY(check_kind(h, KIND_MATHOP, &m));
    CLB_CHECK(nitems >= 0, CLB200_EINVAL, "negative item count");
    if (nitems == 0) return CLB200_OK;   // reference returns 0 items (clMathOp_impl.cc:367-369)
    DeviceGuard g(m->device);
    PortDesc pd;
    pd.nin = 2;
    pd.nout = 1;
    pd.in[0] = a;
    pd.in[1] = b;
    pd.out[0] = c;
    pd.in_bytes[0] = pd.in_bytes[1] = pd.out_bytes[0] = 4 * floats_per_item(m->dtype);
    Bin f{m->op, m->dtype == CLB200_DTYPE_COMPLEX, m->dtype == CLB200_DTYPE_INT};
    int sms = device_sm_count(m->device);
    int fpi = floats_per_item(m->dtype);
    return run_chunked(m, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           m->n_launch++;
                           return launch_map2(di[0], di[1], dout[0], n * fpi, f, sms, st, m);
                       });
}

// ---- secondary ------------------------------------------------------------------
static int unary_launch(Unary *u, const void *in, void *out, long n, cudaStream_t st)
{
    int sms = device_sm_count(u->device);
    u->n_launch++;
    switch (u->ukind) {
    case CLB200_UNARY_LOG10:
        k_log10<<<tile_grid(n / 4, sms), EW_THREADS, 0, st>>>((const float *)in, (float *)out, n,
                                                               (float)((double)u->nv / 3.321928094887362),
                                                               u->kv, tile_wq(u, n / 4, sms, st));
        break;
    case CLB200_UNARY_COMPLEX_TO_MAG:
        k_c2mp<true, false><<<tile_grid(n / 2, sms), EW_THREADS, 0, st>>>((const float2 *)in, (float *)out,
                                                                         nullptr, n, tile_wq(u, n / 2, sms, st));
        break;
    case CLB200_UNARY_COMPLEX_TO_ARG:
        k_c2mp<false, true><<<tile_grid(n / 2, sms), EW_THREADS, 0, st>>>((const float2 *)in, nullptr,
                                                                         (float *)out, n, tile_wq(u, n / 2, sms, st));
        break;
    }
    CLB_CUDA(cudaGetLastError());
    return CLB200_OK;
}

int clb200_unary_create(int kind, int device, float n_value, float k_value, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    CLB_CHECK(kind >= CLB200_UNARY_LOG10 && kind <= CLB200_UNARY_COMPLEX_TO_ARG, CLB200_EINVAL,
              "unknown unary kind %d", kind);
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    Unary *u = new Unary;
    u->kind = KIND_UNARY;
    u->device = device;
    u->init_work_counters();
    u->ukind = kind;
    u->nv = n_value;
    u->kv = k_value;
    *out = u;
    return CLB200_OK;
}

int clb200_unary_launch_device(clb200_handle h, const void *d_in, void *d_out, long nitems,
                               void *stream)
{
    Unary *u;
    CLB_TRY(check_kind(h, KIND_UNARY, &u));
    if (nitems <= 0) return CLB200_OK;
    DeviceGuard g(u->device);
    return unary_launch(u, d_in, d_out, nitems, (cudaStream_t)stream);
}

int clb200_unary_work(clb200_handle h, const void *in, void *out, long nitems)
{
    Unary *u;
    CLB_TRY(check_kind(h, KIND_UNARY, &u));
    if (nitems <= 0) return CLB200_OK;
    DeviceGuard g(u->device);
    PortDesc pd;
    pd.nin = pd.nout = 1;
    pd.in[0] = in;
    pd.out[0] = out;
    pd.in_bytes[0] = (u->ukind == CLB200_UNARY_LOG10) ? 4 : 8;
    pd.out_bytes[0] = 4;
    return run_chunked(u, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           return unary_launch(u, di[0], dout[0], n, st);
                       });
}

static int simple_create(int kind, int device, float nv, float kv, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    Unary *u = new Unary;
    u->kind = kind;
    u->device = device;
    u->ukind = 0;
    u->nv = nv;
    u->kv = kv;
    *out = u;
    return CLB200_OK;
}

int clb200_snr_create(int device, float n_value, float k_value, clb200_handle *out)
{
    return simple_create(KIND_SNR, device, n_value, k_value, out);
}

int clb200_snr_work(clb200_handle h, const float *a, const float *b, float *c, long nitems)
{
    Unary *u;
    CLB_TRY(check_kind(h, KIND_SNR, &u));
    if (nitems <= 0) return CLB200_OK;
    DeviceGuard g(u->device);
    PortDesc pd;
    pd.nin = 2;
    pd.nout = 1;
    pd.in[0] = a;
    pd.in[1] = b;
    pd.out[0] = c;
    pd.in_bytes[0] = pd.in_bytes[1] = pd.out_bytes[0] = 4;
    int sms = device_sm_count(u->device);
    return run_chunked(u, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           u->n_launch++;
                           k_snr<<<item_grid(n, sms), EW_THREADS, 0, st>>>(
                               (const float *)di[0], (const float *)di[1], (float *)dout[0], n, u->nv,
                               u->kv);
                           CLB_CUDA(cudaGetLastError());
                           return CLB200_OK;
                       });
}

int clb200_c2magphase_create(int device, clb200_handle *out)
{
    return simple_create(KIND_C2MAGPHASE, device, 0.f, 0.f, out);
}

int clb200_c2magphase_work(clb200_handle h, const void *in, float *mag, float *phase, long nitems)
{
    Unary *u;
    CLB_TRY(check_kind(h, KIND_C2MAGPHASE, &u));
    if (nitems <= 0) return CLB200_OK;
    DeviceGuard g(u->device);
    PortDesc pd;
    pd.nin = 1;
    pd.nout = 2;
    pd.in[0] = in;
    pd.out[0] = mag;
    pd.out[1] = phase;
    pd.in_bytes[0] = 8;
    pd.out_bytes[0] = pd.out_bytes[1] = 4;
    int sms = device_sm_count(u->device);
    return run_chunked(u, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           u->n_launch++;
                           k_c2mp<true, true><<<tile_grid(n / 2, sms), EW_THREADS, 0, st>>>(
                               (const float2 *)di[0], (float *)dout[0], (float *)dout[1], n, tile_wq(u, n / 2, sms, st));
                           CLB_CUDA(cudaGetLastError());
                           return CLB200_OK;
                       });
}

int clb200_magphase2c_create(int device, clb200_handle *out)
{
    return simple_create(KIND_MAGPHASE2C, device, 0.f, 0.f, out);
}

int clb200_magphase2c_work(clb200_handle h, const float *mag, const float *phase, void *out,
                           long nitems)
{
    Unary *u;
    CLB_TRY(check_kind(h, KIND_MAGPHASE2C, &u));
    if (nitems <= 0) return CLB200_OK;
    DeviceGuard g(u->device);
    PortDesc pd;
    pd.nin = 2;
    pd.nout = 1;
    pd.in[0] = mag;
    pd.in[1] = phase;
    pd.out[0] = out;
    pd.in_bytes[0] = pd.in_bytes[1] = 4;
    pd.out_bytes[0] = 8;
    int sms = device_sm_count(u->device);
    return run_chunked(u, pd, nitems, chunk_for(pd),
                       [&](const void **di, void **dout, long n, cudaStream_t st, long *) {
                           u->n_launch++;
                           k_mp2c<<<item_grid(n, sms), EW_THREADS, 0, st>>>(
                               (const float *)di[0], (const float *)di[1], (float2 *)dout[0], n);
                           CLB_CUDA(cudaGetLastError());
                           return CLB200_OK;
                       });
}

} // extern "C"
