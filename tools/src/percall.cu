// Per-call latency through the C ABI (no python): 8192-sample pageable buffers, like the
// reference harness (test_clenabled.cc:1237-1251).  Also prints raw CUDA floor numbers.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/clenabled_b200.h"
__global__ void k_empty() {}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 8192, iters = 2000;
    std::vector<float> a(2 * n, 1.0f), b(2 * n, 0.5f), c(2 * n);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    k_empty<<<1, 32, 0, st>>>();
    cudaStreamSynchronize(st);
    double t0 = now();
    for (int i = 0; i < iters; i++) { k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }
    printf("empty kernel launch+sync      %6.2f us\n", (now() - t0) / iters * 1e6);
    cudaPointerAttributes at;
    t0 = now();
    for (int i = 0; i < iters; i++) cudaPointerGetAttributes(&at, a.data());
    printf("cudaPointerGetAttributes      %6.2f us\n", (now() - t0) / iters * 1e6);
    clb200_handle h;
    if (clb200_mathconst_create(CLB200_DTYPE_COMPLEX, 0, 2.0f, CLB200_OP_MULTIPLY, &h)) { printf("%s\n", clb200_last_error()); return 1; }
    clb200_mathconst_work(h, a.data(), c.data(), n);
    t0 = now();
    for (int i = 0; i < iters; i++) clb200_mathconst_work(h, a.data(), c.data(), n);
    double dt = (now() - t0) / iters;
    printf("clMultiplyConst work(%d)     %6.2f us  %8.1f Msamples/s\n", n, dt * 1e6, n / dt / 1e6);
    clb200_handle f;
    clb200_fft_create(8192, -1, nullptr, 0, CLB200_DTYPE_COMPLEX, 0, 0, &f);
    clb200_fft_work(f, a.data(), c.data(), n / 8192);
    t0 = now();
    for (int i = 0; i < iters; i++) clb200_fft_work(f, a.data(), c.data(), n / 8192);
    dt = (now() - t0) / iters;
    printf("clFFT 8192 work(%d)          %6.2f us  %8.1f Msamples/s\n", n, dt * 1e6, n / dt / 1e6);
    clb200_destroy(h);
    clb200_destroy(f);
    return 0;
}
