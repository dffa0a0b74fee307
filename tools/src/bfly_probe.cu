// bfly_probe.cu -- what the FP32 pipe of one SM sustains on the FFT filter's own instruction mix, without any memory
// traffic: every thread keeps 32 complex values in registers and runs the radix-32 butterfly of fft_device.cuh
// (FADD2 adds, FMUL/FFMA/FADD twiddle rotations) back to back, optionally followed by 31 complex multiplies by a
// loop-invariant factor (the inter-pass twiddle / tap-spectrum products).  One-warp CTAs, 1..4 warps per scheduler.
// lane-ops per iteration: butterfly 456 (320 add + 136 rotation), products 31 * 4 = 124.
#include "../../gr_clenabled_b200/csrc/fft_device.cuh"
#include <cstdio>
using namespace clb200::fftdev;

// UNROLL copies of the loop body (instruction footprint UNROLL x ~8.5 KB) and a per-warp start delay, so that the
// warps of a scheduler sit at different places of a loop that does not fit the instruction caches next to the SM
template <int UNROLL>
__global__ void __launch_bounds__(32) k_bfly_code(float2 *out, int iters, float2 w, int desync)
{
    float2 x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f - blockIdx.x);
    if (desync) {
        const long long t0 = clock64();
        const long long wait = (long long)((blockIdx.x * 2654435761u) % 20000u);
        while (clock64() - t0 < wait) { }
    }
    for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
        for (int c = 0; c < UNROLL; c++) {
#pragma unroll
            for (int r = 1; r < 32; r++) x[r] = cmul(x[r], w);
            dft_dif<32, 0>(x);
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 32; i++) { s.x += x[i].x; s.y += x[i].y; }
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

template <int UNROLL> void run_code(float2 *d, int sms, int warps_per_sm, int desync)
{
    const int iters = 4000;
    const float2 w = make_float2(0.03125f, 0.0001f);
    k_bfly_code<UNROLL><<<sms * warps_per_sm, 32>>>(d, 8, w, desync);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_bfly_code<UNROLL><<<sms * warps_per_sm, 32>>>(d, iters, w, desync);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double laneops = (double)sms * warps_per_sm * 32 * iters * 580.0;
    printf("loop body x%d (%s)  %2d warps/SM  %8.3f ms  %6.2f T lane-ops/s\n", UNROLL, desync ? "staggered" : "lockstep ", warps_per_sm, ms,
           laneops / ms / 1e9);
}

template <int TW>
__global__ void __launch_bounds__(32) k_bfly(float2 *out, int iters, float2 w)
{
    float2 x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f - blockIdx.x);
    for (int it = 0; it < iters; it++) {
        if (TW) {
#pragma unroll
            for (int r = 1; r < 32; r++) x[r] = cmul(x[r], w);
        }
        dft_dif<32, 0>(x);
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 32; i++) { s.x += x[i].x; s.y += x[i].y; }
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

template <int TW> void run(float2 *d, int sms, int warps_per_sm)
{
    const int iters = 4000;
    const float2 w = make_float2(0.03125f, 0.0001f);
    k_bfly<TW><<<sms * warps_per_sm, 32>>>(d, 10, w);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_bfly<TW><<<sms * warps_per_sm, 32>>>(d, iters, w);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double laneops = (double)sms * warps_per_sm * 32 * iters * (456 + (TW ? 124 : 0));
    printf("%s  %2d warps/SM  %8.3f ms  %6.2f T lane-ops/s\n", TW ? "butterfly + 31 products" : "butterfly only         ",
           warps_per_sm, ms, laneops / ms / 1e9);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float2 *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 16 * 32 * 8);
    printf("packed adds (CLB_F32X2) = %d\n", CLB_F32X2);
    for (int w : {4, 8, 12, 16}) run<0>(d, p.multiProcessorCount, w);
    for (int w : {4, 8, 12, 16}) run<1>(d, p.multiProcessorCount, w);
    for (int ds : {0, 1}) {
        run_code<1>(d, p.multiProcessorCount, 12, ds);
        run_code<2>(d, p.multiProcessorCount, 12, ds);
        run_code<4>(d, p.multiProcessorCount, 12, ds);
        run_code<8>(d, p.multiProcessorCount, 12, ds);
        run_code<16>(d, p.multiProcessorCount, 12, ds);
    }
    return 0;
}
