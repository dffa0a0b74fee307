// mailbox_probe.cu -- can a RESIDENT kernel (one CTA polling a mailbox in pinned host memory) serve scheduler-sized
// calls faster than launch + synchronise?  Round trip of a 64 KiB multiply-by-constant: pageable -> pinned memcpy,
// post, kernel reads / writes the pinned buffers over PCIe, completion flag, pinned -> pageable memcpy.
// The kernel leaves by itself after an idle time-out and after a hard lifetime cap.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <atomic>
#include <immintrin.h>
#include <thread>

struct Mailbox {
    volatile unsigned long long req, done;
    volatile unsigned int alive, quit;
    volatile long nvec;
    volatile float k;
    const float4 *volatile in;
    float4 *volatile out;
    volatile unsigned long long ns_params, ns_read, ns_write, ns_fence;
    volatile int variant;      // bit 0: plain loads, bit 1: plain stores, bit 2: fence by every thread
};

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(1024) k_resident(Mailbox *mb, unsigned long long idle_ns, unsigned long long life_ns)
{
    __shared__ unsigned long long s_seq;
    __shared__ int s_cmd;
    unsigned long long last = mb->done;
    const unsigned long long born = gtime();
    unsigned long long idle0 = born;
    for (;;) {
        if (threadIdx.x == 0) {
            for (;;) {
                const unsigned long long seq = mb->req;
                if (seq != last) { s_seq = seq; s_cmd = 1; break; }
                const unsigned long long now = gtime();
                if (mb->quit || now - idle0 > idle_ns || now - born > life_ns) { s_cmd = 2; break; }
            }
        }
        __syncthreads();
        if (s_cmd == 2) break;
        const unsigned long long t0 = gtime();
        const long nvec = mb->nvec;
        const float k = mb->k;
        const float4 *in = mb->in;
        float4 *out = mb->out;
        const int variant = mb->variant;
        const unsigned long long t1 = gtime();
        float4 v[4];
        const long i = threadIdx.x;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u * 1024 < nvec) v[u] = (variant & 1) ? __ldcs(in + i + u * 1024) : __ldcv(in + i + u * 1024);
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) sum += v[u].x;
        if (sum == 12345.678f) mb->k = 0.f;        // keep the loads before t2
        __syncthreads();
        const unsigned long long t2 = gtime();
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u * 1024 < nvec) {
                const float4 o = make_float4(v[u].x * k, v[u].y * k, v[u].z * k, v[u].w * k);
                if (variant & 2) out[i + u * 1024] = o;
                else __stwt(out + i + u * 1024, o);
            }
        __syncthreads();
        const unsigned long long t3 = gtime();
        if (variant & 4) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned long long t4 = gtime();
            mb->ns_params = t1 - t0; mb->ns_read = t2 - t1; mb->ns_write = t3 - t2; mb->ns_fence = t4 - t3;
            last = s_seq;
            mb->done = last;
            idle0 = gtime();
        }
    }
    if (threadIdx.x == 0) {
        mb->alive = 0;
        __threadfence_system();
    }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    const int n = 8192;                         // complex samples
    const size_t bytes = (size_t)n * 8;
    std::vector<float> a(2 * n, 1.5f), c(2 * n, 0.f);
    Mailbox *mb;
    float *pin_in, *pin_out;
    cudaHostAlloc(&mb, sizeof(Mailbox), cudaHostAllocMapped);
    cudaHostAlloc(&pin_in, bytes, cudaHostAllocMapped);
    cudaHostAlloc(&pin_out, bytes, cudaHostAllocMapped);
    memset((void *)mb, 0, sizeof(Mailbox));
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    mb->in = (const float4 *)pin_in;
    mb->out = (float4 *)pin_out;
    mb->nvec = bytes / 16;
    mb->k = 2.0f;
    unsigned long long seq = 0;
    long relaunches = 0;
    auto launch = [&]() {
        mb->alive = 1;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        k_resident<<<1, 1024, 0, st>>>(mb, 2000000ull /* 2 ms idle */, 3000000000ull /* 3 s cap */);
        relaunches++;
    };
    auto call = [&](bool copy) -> bool {
        if (copy) memcpy(pin_in, a.data(), bytes);
        std::atomic_thread_fence(std::memory_order_release);
        mb->req = ++seq;
        const double t0 = now();
        while (mb->done != seq) {
            if (!mb->alive) launch();
            _mm_pause();
            if (now() - t0 > 0.5) { printf("timeout\n"); return false; }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if (copy) memcpy(c.data(), pin_out, bytes);
        return true;
    };
    launch();
    for (int w = 0; w < 50; w++) if (!call(true)) return 1;
    for (int variant = 0; variant < 8; variant++) {
        mb->variant = variant;
        for (int mode = 0; mode < 2; mode++) {
            const int iters = 3000;
            const double t0 = now();
            for (int i = 0; i < iters; i++) if (!call(mode == 0)) return 1;
            const double dt = (now() - t0) / iters;
            printf("variant %d (%s loads, %s stores, fence by %s), %s: %6.2f us per call | on the GPU: params %llu ns, read %llu, write %llu, fence %llu\n",
                   variant, variant & 1 ? "plain" : "ld.cv", variant & 2 ? "plain" : "st.wt", variant & 4 ? "all" : "one",
                   mode == 0 ? "pageable" : "pinned  ", dt * 1e6, mb->ns_params, mb->ns_read, mb->ns_write, mb->ns_fence);
        }
    }
    mb->variant = 0;
    call(true);
    bool ok = true;
    for (int i = 0; i < 2 * n; i++) ok = ok && c[i] == 3.0f;
    printf("result %s, kernel launches %ld\n", ok ? "correct" : "WRONG", relaunches);
    // a call after the idle time-out: the kernel has left and is relaunched
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    double t0 = now();
    call(true);
    printf("first call after 20 ms idle: %.1f us (relaunch), launches %ld\n", (now() - t0) * 1e6, relaunches);
    mb->quit = 1;
    cudaStreamSynchronize(st);
    printf("alive after quit: %u\n", mb->alive);
    // the launch + synchronise way on the same buffers, for reference
    return 0;
}
