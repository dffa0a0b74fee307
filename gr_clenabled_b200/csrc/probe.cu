// probe.cu -- measured FP32 pipe peaks of the device, for the compute-side roofline of the FP32-bound kernels
// (FFT filter, time-domain FIR; SURVEY 8d row 3 asks for the box's own FFMA peak, not a data-sheet figure).
#include "common.cuh"

using namespace clb200;

namespace {

// 16 independent accumulator chains per thread: enough ILP to keep the FMA pipe issuing every cycle
template <int MODE>      // 0: FFMA (2 flop per lane-op), 1: FADD (1 flop), 2: packed add.f32x2 (2 lanes per instruction)
__global__ void __launch_bounds__(256) k_probe_fp32(float *out, int iters, float b, float c)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = __fadd_rn(a[i], c);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long x, y, z;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(c), "f"(b));
                    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(z));
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345.678f) out[0] = s;           // keeps the chains alive
}

template <int MODE>
int run_probe(int device, double *lane_ops_per_s)
{
    const int sms = device_sm_count(device);
    const int grid = sms * 8, iters = 4096;
    float *d = nullptr;
    CLB_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CLB_CUDA(cudaEventCreate(&e0));
    CLB_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CLB_CUDA(cudaEventRecord(e0, 0));
        k_probe_fp32<MODE><<<grid, 256>>>(d, iters, 1.0000001f, 1e-9f);
        CLB_CUDA(cudaEventRecord(e1, 0));
        CLB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)grid * 256 * iters * 4 * 16;        // per-lane operations
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *lane_ops_per_s = best;
    return CLB200_OK;
}

} // namespace

extern "C" int clb200_probe_fp32(int device, double *ffma_tflops, double *fadd_tflops, double *fadd2_tflops)
{
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    DeviceGuard g(device);
    double v = 0;
    if (ffma_tflops) {
        CLB_TRY(run_probe<0>(device, &v));
        *ffma_tflops = 2.0 * v / 1e12;
    }
    if (fadd_tflops) {
        CLB_TRY(run_probe<1>(device, &v));
        *fadd_tflops = v / 1e12;
    }
    if (fadd2_tflops) {
        CLB_TRY(run_probe<2>(device, &v));
        *fadd2_tflops = v / 1e12;
    }
    return CLB200_OK;
}
