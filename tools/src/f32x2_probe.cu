// f32x2_probe.cu -- per-SM throughput of scalar FADD/FFMA vs packed add/fma.f32x2 (FADD2/FFMA2) on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, int iters)
{
    float2 v[8];
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 w = make_float2(1.0001f, 0.9999f);
    unsigned long long ww = *reinterpret_cast<const unsigned long long *>(&w);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { v[i].x += w.x; v[i].y += w.y; }
            if (MODE == 1) { unsigned long long &x = *reinterpret_cast<unsigned long long *>(&v[i]);
                             asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(ww)); }
            if (MODE == 2) { v[i].x = fmaf(v[i].x, w.x, w.y); v[i].y = fmaf(v[i].y, w.y, w.x); }
            if (MODE == 3) { unsigned long long &x = *reinterpret_cast<unsigned long long *>(&v[i]);
                             asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x) : "l"(ww)); }
        }
    }
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; i++) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, float2 *d, int sms)
{
    const int iters = 20000;
    k<MODE><<<sms * 4, 256>>>(d, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<sms * 4, 256>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double complex_ops = (double)sms * 4 * 256 * 8 * iters;      // one op on a (re, im) pair
    printf("%-10s %8.3f ms  %7.1f G complex-ops/s  (%.1f per SM per ns)\n", name, ms, complex_ops / ms / 1e6, complex_ops / ms / 1e6 / sms);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float2 *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 256 * 8);
    run<0>("2xFADD", d, p.multiProcessorCount);
    run<1>("FADD2", d, p.multiProcessorCount);
    run<2>("2xFFMA", d, p.multiProcessorCount);
    run<3>("FFMA2", d, p.multiProcessorCount);
    return 0;
}
