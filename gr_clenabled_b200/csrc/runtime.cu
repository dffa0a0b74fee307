// runtime.cu -- device selection, buffers, error state, handle lifetime.
// Replaces GRCLBase::InitOpenCL / cleanup (lib/GRCLBase.cpp:17-369, :423-483).
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdlib>

namespace clb200 {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count(int device)
{
    static int cache[64];
    if (device < 0 || device >= 64) return 148;
    if (cache[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
            v <= 0)
            v = 148;
        cache[device] = v;
    }
    return cache[device];
}

size_t small_call_bytes()
{
    static long v = -1;
    if (v < 0) {
        const char *e = getenv("CLB200_SMALL_KB");
        long kb = e ? atol(e) : 1024;
        v = (kb >= 0 && kb <= (1 << 20) ? kb : 1024) << 10;
    }
    return (size_t)v;
}

size_t chunk_target_bytes()
{
    static size_t v = 0;
    if (v == 0) {
        const char *e = getenv("CLB200_CHUNK_MB");
        long mb = e ? atol(e) : 32;       // 32 MiB: best PCIe overlap measured on B200 (tools/e2e_fft.py)
        v = (size_t)(mb >= 1 && mb <= 1024 ? mb : 32) << 20;
    }
    return v;
}

int Buf::reserve(size_t bytes)
{
    if (bytes <= cap) return CLB200_OK;
    release();
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("%s of %zu bytes failed: %s", host ? "cudaHostAlloc" : "cudaMalloc", want,
                  cudaGetErrorString(e));
        cudaGetLastError();
        return CLB200_ENOMEM;
    }
    cap = want;
    return CLB200_OK;
}

void Buf::release()
{
    if (p) {
        if (host) cudaFreeHost(p);
        else cudaFree(p);
    }
    p = nullptr;
    cap = 0;
}

// Scheduler-sized calls come thousands of times per second on the same few (pageable) buffers, and
// cudaPointerGetAttributes costs ~0.5 us per port: pages that were seen to be PAGEABLE are remembered in a small
// direct-mapped table (a stale "pageable" only costs the staging copy; page-locked pointers are always re-checked,
// so a buffer that was freed or unregistered is never handed to a kernel).  register / unregister clear the table.
static std::atomic<uintptr_t> g_pageable[256];

void pinned_cache_clear()
{
    for (auto &e : g_pageable) e.store(0, std::memory_order_relaxed);
}

bool is_pinned(const void *p)
{
    const uintptr_t page = (uintptr_t)p >> 12;
    std::atomic<uintptr_t> &slot = g_pageable[page & 255];
    if (slot.load(std::memory_order_relaxed) == page) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a.type == cudaMemoryTypeHost) return true;
    slot.store(page, std::memory_order_relaxed);
    return false;
}

} // namespace clb200

using namespace clb200;

void clb200_block::set_info(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    info = buf;
}

int clb200_block::init_slots()
{
    if (slots_ready) return CLB200_OK;
    for (int i = 0; i < NSLOT; i++) {
        Slot &s = slot[i];
        for (int k = 0; k < MAXPORT; k++) {
            s.pin_in[k].host = true;
            s.pin_out[k].host = true;
        }
        CLB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CLB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    slots_ready = true;
    return CLB200_OK;
}

// the records are allocated and zeroed when the handle is created (a cudaMemset at launch time would synchronise
// with the caller's streams, and is not allowed while one of them is being captured into a graph)
bool clb200_block::init_work_counters()
{
    if (work_ctr_buf.p) return true;
    DeviceGuard g(device);
    if (work_ctr_buf.reserve(16 * WORK_CTRS) != CLB200_OK) return false;
    if (cudaMemset(work_ctr_buf.p, 0, 16 * WORK_CTRS) != cudaSuccess) {
        cudaGetLastError();
        work_ctr_buf.release();
        return false;
    }
    return true;
}

unsigned long long *clb200_block::work_counter(cudaStream_t st)
{
    const char *e = getenv("CLB200_STATIC_TILES");      // A/B switch, read per launch
    if (e && atoi(e) != 0) return nullptr;
    if (!work_ctr_buf.p && !init_work_counters()) return nullptr;
    int i = 0;
    for (; i < work_ctr_used; i++)
        if (work_ctr_stream[i] == st) break;
    if (i == work_ctr_used) {
        if (work_ctr_used == WORK_CTRS) return nullptr;
        work_ctr_stream[work_ctr_used++] = st;
    }
    return reinterpret_cast<unsigned long long *>(work_ctr_buf.p) + 2 * i;
}

clb200_block::~clb200_block()
{
    DeviceGuard g(device);
    work_ctr_buf.release();
    for (int i = 0; i < NSLOT; i++) {
        Slot &s = slot[i];
        if (s.stream) cudaStreamSynchronize(s.stream);
        for (int k = 0; k < MAXPORT; k++) {
            s.pin_in[k].release();
            s.pin_out[k].release();
            s.dev_in[k].release();
            s.dev_out[k].release();
        }
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
}

extern "C" {

const char *clb200_version(void) { return "clenabled_b200 0.1 (sm_100a)"; }

const char *clb200_last_error(void) { return g_err; }

int clb200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        cudaGetLastError();
        return 0;
    }
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return CLB200_ECUDA;
    }
    return n;
}

int clb200_device_name(int device, char *buf, int buflen)
{
    CLB_CHECK(buf && buflen > 0, CLB200_EINVAL, "bad buffer");
    cudaDeviceProp p;
    CLB_CUDA(cudaGetDeviceProperties(&p, device));
    snprintf(buf, buflen, "%s (sm_%d%d, %d SMs, %.1f GB)", p.name, p.major, p.minor,
             p.multiProcessorCount, (double)p.totalGlobalMem / 1e9);
    return CLB200_OK;
}

int clb200_device_sm_count(int device) { return device_sm_count(device); }

int clb200_select_device(int platform_type, int dev_selector, int platform_id, int dev_id)
{
    (void)platform_id;
    // OCLTYPE_GPU=1, ACCELERATOR=2, CPU=3, ANY=4 (GRCLBase.h:64-67): there is no
    // CPU device behind this library; a flowgraph that asks for one gets the GPU.
    CLB_CHECK(platform_type >= 1 && platform_type <= 4, CLB200_EINVAL,
              "openCLPlatformType %d is not one of 1..4", platform_type);
    int n = clb200_device_count();
    if (n < 0) return n;
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    int dev = (dev_selector == 2) ? dev_id : 0;   // OCLDEVICESELECTOR_SPECIFIC=2 (GRCLBase.h:69-70)
    CLB_CHECK(dev >= 0 && dev < n, CLB200_EINVAL, "device %d requested, %d present", dev, n);
    return dev;
}

int clb200_register_host_buffer(void *ptr, size_t bytes)
{
    CLB_CHECK(ptr != nullptr && bytes > 0, CLB200_EINVAL, "bad host range");
    CLB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    pinned_cache_clear();
    return CLB200_OK;
}

int clb200_unregister_host_buffer(void *ptr)
{
    CLB_CHECK(ptr != nullptr, CLB200_EINVAL, "null pointer");
    CLB_CUDA(cudaHostUnregister(ptr));
    pinned_cache_clear();
    return CLB200_OK;
}

int clb200_destroy(clb200_handle h)
{
    if (!h) return CLB200_OK;
    delete h;
    return CLB200_OK;
}

int clb200_describe(clb200_handle h, char *buf, int buflen)
{
    CLB_CHECK(h != nullptr && buf != nullptr && buflen > 0, CLB200_EINVAL, "bad arguments");
    snprintf(buf, buflen, "device %d, %d SMs: %s", h->device, device_sm_count(h->device),
             h->info.empty() ? "(no description)" : h->info.c_str());
    return CLB200_OK;
}

int clb200_set_debug(clb200_handle h, int on)
{
    CLB_CHECK(h != nullptr, CLB200_EINVAL, "null handle");
    h->debug = on ? 1 : 0;
    return CLB200_OK;
}

int clb200_get_counters(clb200_handle h, uint64_t *h2d, uint64_t *d2h, uint64_t *launches)
{
    CLB_CHECK(h != nullptr, CLB200_EINVAL, "null handle");
    if (h2d) *h2d = h->n_h2d;
    if (d2h) *d2h = h->n_d2h;
    if (launches) *launches = h->n_launch;
    return CLB200_OK;
}


// ---- device memory that other ranks of the box can map (fused X-engine gather) ----
int clb200_mem_alloc(int device, size_t bytes, void **dptr)
{
    CLB_CHECK(dptr != nullptr && bytes > 0, CLB200_EINVAL, "bad arguments");
    DeviceGuard g(device);
    CLB_CUDA(cudaMalloc(dptr, bytes));
    CLB_CUDA(cudaMemset(*dptr, 0, bytes));
    return CLB200_OK;
}
int clb200_mem_free(int device, void *dptr)
{
    DeviceGuard g(device);
    CLB_CUDA(cudaFree(dptr));
    return CLB200_OK;
}
int clb200_mem_copy_to_host(int device, const void *dptr, void *host, size_t bytes)
{
    DeviceGuard g(device);
    CLB_CUDA(cudaMemcpy(host, dptr, bytes, cudaMemcpyDeviceToHost));
    return CLB200_OK;
}
int clb200_ipc_export(int device, void *dptr, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "interprocess handle is 64 bytes");
    CLB_CHECK(dptr && handle64, CLB200_EINVAL, "bad arguments");
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    CLB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return CLB200_OK;
}
int clb200_ipc_open(int device, const void *handle64, void **dptr)
{
    CLB_CHECK(dptr && handle64, CLB200_EINVAL, "bad arguments");
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CLB_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CLB200_OK;
}
int clb200_ipc_close(int device, void *dptr)
{
    DeviceGuard g(device);
    CLB_CUDA(cudaIpcCloseMemHandle(dptr));
    return CLB200_OK;
}

} // extern "C"
