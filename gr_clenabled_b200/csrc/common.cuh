// common.cuh -- shared host-side plumbing of libclenabled_b200.so
//
// Replaces the role of GRCLBase (include/clenabled/GRCLBase.h:77-143,
// lib/GRCLBase.cpp): device selection, queue (here: CUDA streams), buffers and
// error reporting.  Where GRCLBase prints and exit(0)s on a runtime failure
// (GRCLBase.cpp:239-257,402-410), every failure here becomes an int status +
// a thread-local message.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>

#include "../../include/clenabled_b200.h"

namespace clb200 {

void set_error(const char *fmt, ...);

#define CLB_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ::clb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,       \
                                cudaGetErrorString(_e));                            \
            return CLB200_ECUDA;                                                    \
        }                                                                           \
    } while (0)

#define CLB_CHECK(cond, code, ...)                                                  \
    do {                                                                            \
        if (!(cond)) {                                                              \
            ::clb200::set_error(__VA_ARGS__);                                       \
            return (code);                                                          \
        }                                                                           \
    } while (0)

#define CLB_TRY(expr)                                                               \
    do {                                                                            \
        int _rc = (expr);                                                           \
        if (_rc != CLB200_OK) return _rc;                                           \
    } while (0)

// NVTX range for the host-path phases (H2D / kernels / D2H); a no-op unless a profiler is attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

enum BlockKind {
    KIND_MATHCONST = 1, KIND_MATHOP, KIND_UNARY, KIND_SNR, KIND_C2MAGPHASE, KIND_MAGPHASE2C,
    KIND_FFT, KIND_FILTER, KIND_PFB, KIND_XENGINE, KIND_XCFFT, KIND_XCORR, KIND_CFILTER, KIND_QUADDEMOD, KIND_SIGSOURCE
};

int device_sm_count(int device);

// A growable device or pinned-host buffer.
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    bool host = false;
    int reserve(size_t bytes);   // grows (never shrinks); contents are NOT preserved
    void release();
};

constexpr int NSLOT = 3;   // H2D of chunk i+1 | kernel of chunk i | D2H of chunk i-1
constexpr int MAXPORT = 4;

// One pipeline slot: its own stream, pinned staging and device buffers.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    Buf pin_in[MAXPORT], pin_out[MAXPORT], dev_in[MAXPORT], dev_out[MAXPORT];
    bool busy = false;
    // pending drain description
    long first = 0, n = 0;
};

} // namespace clb200

// The opaque handle.  Every block type derives from it.
struct clb200_block {
    int kind = 0;
    int device = 0;
    std::mutex mtx;                 // guards setters vs work (reference: d_mutex)
    clb200::Slot slot[clb200::NSLOT];
    bool slots_ready = false;
    uint64_t n_h2d = 0, n_d2h = 0, n_launch = 0;
    std::string info;               // kernel variant and launch geometry, filled by create (clb200_describe)
    int debug = 0;                  // setDebug of the reference factories (clb200_set_debug)
    void set_info(const char *fmt, ...);
    virtual ~clb200_block();
    int init_slots();
    // Work counters of the persistent kernels (tile_fetch / tile_finish below): one 16-byte record {next tile,
    // finished CTAs} per launching stream -- launches on one stream are serialised and a kernel's last CTA leaves
    // its record at zero.  nullptr (more than WORK_CTRS streams, allocation failure, CLB200_STATIC_TILES=1) makes
    // the kernel fall back to static striding.
    static constexpr int WORK_CTRS = 8;
    clb200::Buf work_ctr_buf;
    cudaStream_t work_ctr_stream[WORK_CTRS] = {};
    int work_ctr_used = 0;
    bool init_work_counters();
    unsigned long long *work_counter(cudaStream_t st);
};

namespace clb200 {

bool is_pinned(const void *p);
void pinned_cache_clear();
size_t small_call_bytes();

struct PortDesc {
    const void *in[MAXPORT] = {nullptr, nullptr, nullptr, nullptr};
    void *out[MAXPORT] = {nullptr, nullptr, nullptr, nullptr};
    size_t in_bytes[MAXPORT] = {0, 0, 0, 0};    // bytes per item on that port
    long in_extra[MAXPORT] = {0, 0, 0, 0};      // bytes read past (or short of, if <0) each chunk: overlap
    size_t out_bytes[MAXPORT] = {0, 0, 0, 0};
    int nin = 0, nout = 0;
};

// Stream `nitems` items through the handle's slots in chunks: for each chunk
//   host->pinned (only if the caller's buffer is pageable) -> H2D -> launch -> D2H
// `launch(d_in[], d_out[], n, stream, &n_out)` enqueues the kernels for n input
// items and reports how many output items they produce (n_out <= n; = n unless
// the block decimates).  All outputs are in the caller's buffers when this
// returns; *total_out (optional) is the number of output items written.
template <class Launch>
int run_chunked(clb200_block *b, const PortDesc &pd, long nitems, long chunk_items, Launch launch,
                long *total_out = nullptr)
{
    if (total_out) *total_out = 0;
    if (nitems <= 0) return CLB200_OK;
    const auto t_start = std::chrono::steady_clock::now();
    auto debug_line = [&](const char *route, long chunks, long n_out) {
        if (!b->debug) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_start).count();
        fprintf(stderr, "clenabled_b200[debug] kind %d: work(%ld items) -> %ld items, %s, %ld chunk(s), %.1f us\n", b->kind,
                nitems, n_out, route, chunks, us);
    };
    CLB_TRY(b->init_slots());
    bool pin_i[MAXPORT], pin_o[MAXPORT];
    for (int k = 0; k < pd.nin; k++) pin_i[k] = is_pinned(pd.in[k]);
    for (int k = 0; k < pd.nout; k++) pin_o[k] = is_pinned(pd.out[k]);
    if (chunk_items < 1) chunk_items = 1;
    if (chunk_items > nitems) chunk_items = nitems;

    // Scheduler-sized calls (the reference's 8192-sample buffers are 64 KiB): the copy
    // engines' latency would dominate, so the kernel reads and writes the pinned staging
    // buffers (or the caller's own pinned buffers) directly over PCIe -- one launch, no
    // cudaMemcpy, one stream synchronise.
    {
        size_t total = 0;
        for (int k = 0; k < pd.nin; k++) total += (size_t)((long)nitems * (long)pd.in_bytes[k] + pd.in_extra[k]);
        for (int k = 0; k < pd.nout; k++) total += (size_t)nitems * pd.out_bytes[k];
        if (total <= small_call_bytes()) {
            Slot &s = b->slot[0];
            const void *d_in[MAXPORT];
            void *d_out[MAXPORT];
            for (int k = 0; k < pd.nin; k++) {
                size_t bytes = (size_t)((long)nitems * (long)pd.in_bytes[k] + pd.in_extra[k]);
                if (pin_i[k]) {
                    d_in[k] = pd.in[k];
                } else {
                    CLB_TRY(s.pin_in[k].reserve(bytes));
                    memcpy(s.pin_in[k].p, pd.in[k], bytes);
                    d_in[k] = s.pin_in[k].p;
                }
                b->n_h2d += bytes;
            }
            for (int k = 0; k < pd.nout; k++) {
                if (pin_o[k]) {
                    d_out[k] = pd.out[k];
                } else {
                    CLB_TRY(s.pin_out[k].reserve((size_t)nitems * pd.out_bytes[k]));
                    d_out[k] = s.pin_out[k].p;
                }
            }
            long n_out = nitems;
            NvtxRange r("clb200 small call: kernel over pinned host memory");
            CLB_TRY(launch(d_in, d_out, nitems, s.stream, &n_out));
            CLB_CUDA(cudaStreamSynchronize(s.stream));
            for (int k = 0; k < pd.nout; k++) {
                if (!pin_o[k] && n_out > 0) memcpy(pd.out[k], s.pin_out[k].p, (size_t)n_out * pd.out_bytes[k]);
                b->n_d2h += (size_t)n_out * pd.out_bytes[k];
            }
            if (total_out) *total_out = n_out;
            debug_line("zero-copy (kernel reads/writes pinned host memory)", 1, n_out);
            return CLB200_OK;
        }
    }

    // any early (error) return below leaves nothing in flight and no slot marked busy: a later call
    // must never drain a stale slot into a different caller buffer
    struct SlotGuard {
        clb200_block *b;
        bool armed = true;
        ~SlotGuard()
        {
            if (!armed) return;
            for (int i = 0; i < NSLOT; i++) {
                if (b->slot[i].stream) cudaStreamSynchronize(b->slot[i].stream);
                b->slot[i].busy = false;
            }
            cudaGetLastError();
        }
    } guard{b};

    auto drain = [&](Slot &s) -> int {
        if (!s.busy) return CLB200_OK;
        CLB_CUDA(cudaEventSynchronize(s.done));
        for (int k = 0; k < pd.nout; k++)
            if (!pin_o[k] && s.n > 0)
                memcpy((char *)pd.out[k] + (size_t)s.first * pd.out_bytes[k], s.pin_out[k].p,
                       (size_t)s.n * pd.out_bytes[k]);
        s.busy = false;
        return CLB200_OK;
    };

    long c = 0, out_pos = 0;
    for (long first = 0; first < nitems; first += chunk_items, c++) {
        long n = std::min(chunk_items, nitems - first);
        Slot &s = b->slot[c % NSLOT];
        NvtxRange r("clb200 chunk: H2D + kernel + D2H enqueue");
        CLB_TRY(drain(s));
        const void *d_in[MAXPORT];
        void *d_out[MAXPORT];
        for (int k = 0; k < pd.nin; k++) {
            size_t bytes = (size_t)((long)n * (long)pd.in_bytes[k] + pd.in_extra[k]);
            size_t cap = (size_t)((long)chunk_items * (long)pd.in_bytes[k] + std::max(0L, pd.in_extra[k]));
            CLB_TRY(s.dev_in[k].reserve(cap));
            const char *src = (const char *)pd.in[k] + (size_t)first * pd.in_bytes[k];
            if (!pin_i[k]) {
                CLB_TRY(s.pin_in[k].reserve(cap));
                memcpy(s.pin_in[k].p, src, bytes);
                src = (const char *)s.pin_in[k].p;
            }
            CLB_CUDA(cudaMemcpyAsync(s.dev_in[k].p, src, bytes, cudaMemcpyHostToDevice, s.stream));
            b->n_h2d += bytes;
            d_in[k] = s.dev_in[k].p;
        }
        for (int k = 0; k < pd.nout; k++) {
            CLB_TRY(s.dev_out[k].reserve((size_t)chunk_items * pd.out_bytes[k]));
            d_out[k] = s.dev_out[k].p;
        }
        long n_out = n;
        CLB_TRY(launch(d_in, d_out, n, s.stream, &n_out));
        for (int k = 0; k < pd.nout && n_out > 0; k++) {
            size_t bytes = (size_t)n_out * pd.out_bytes[k];
            char *dst;
            if (pin_o[k]) {
                dst = (char *)pd.out[k] + (size_t)out_pos * pd.out_bytes[k];
            } else {
                CLB_TRY(s.pin_out[k].reserve((size_t)chunk_items * pd.out_bytes[k]));
                dst = (char *)s.pin_out[k].p;
            }
            CLB_CUDA(cudaMemcpyAsync(dst, s.dev_out[k].p, bytes, cudaMemcpyDeviceToHost, s.stream));
            b->n_d2h += bytes;
        }
        CLB_CUDA(cudaEventRecord(s.done, s.stream));
        s.busy = true;
        s.first = out_pos;
        s.n = n_out;
        out_pos += n_out;
    }
    // drain in issue order
    for (long i = 0; i < NSLOT; i++) CLB_TRY(drain(b->slot[(c + i) % NSLOT]));
    guard.armed = false;
    if (total_out) *total_out = out_pos;
    debug_line("copy engines, 3 slots", c, out_pos);
    return CLB200_OK;
}

// pick a chunk (in items) whose largest port moves about `target` bytes
size_t small_call_bytes();       // calls moving at most this much take the zero-copy path (CLB200_SMALL_KB)
size_t chunk_target_bytes();     // default 32 MiB, CLB200_CHUNK_MB overrides (tuning)

inline long chunk_for(const PortDesc &pd, size_t target = 0)
{
    if (target == 0) target = chunk_target_bytes();
    size_t mx = 1;
    for (int k = 0; k < pd.nin; k++) mx = std::max(mx, pd.in_bytes[k]);
    for (int k = 0; k < pd.nout; k++) mx = std::max(mx, pd.out_bytes[k]);
    return (long)std::max<size_t>(1, target / mx);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
int check_kind(clb200_handle h, int kind, T **out)
{
    CLB_CHECK(h != nullptr, CLB200_EINVAL, "null handle");
    CLB_CHECK(h->kind == kind, CLB200_EINVAL, "handle is of kind %d, call needs kind %d", h->kind,
              kind);
    *out = static_cast<T *>(h);
    return CLB200_OK;
}

#ifdef __CUDACC__
// Dynamic tile assignment for persistent kernels.  Static striding (tile += gridDim.x) lets the hardware's unfair
// warp scheduling decide when each CTA finishes: ncu shows 10.1 of 12 resident warps active on average in the FFT
// filter and 14.4 of 16 in the 8192-point FFT -- the early finishers leave their SM under-occupied for the rest of
// the launch.  A CTA's first tile is blockIdx.x, every further one comes from the counter (fetched one tile ahead,
// so the atomic's latency hides behind the tile's work); the last CTA to finish zeroes the record.
__device__ __forceinline__ long tile_fetch(unsigned long long *rec)
{
    return (long)gridDim.x + (long)atomicAdd(rec, 1ULL);
}
__device__ __forceinline__ void tile_finish(unsigned long long *rec)      // one thread per CTA, after its last fetch
{
    __threadfence();
    if (atomicAdd(rec + 1, 1ULL) == (unsigned long long)gridDim.x - 1) {
        rec[0] = 0;
        rec[1] = 0;
    }
}
#endif

inline int grid_for(long work_ctas, int sms, int per_sm)
{
    long g = std::min<long>(work_ctas, (long)sms * per_sm);
    return (int)std::max<long>(1, g);
}

} // namespace clb200
