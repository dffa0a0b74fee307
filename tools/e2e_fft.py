"""e2e clFFT through clb200_fft_work with pinned host buffers, for chunk-size tuning."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
N, nvec = 8192, 8192
hx = torch.empty(N * nvec * 2, dtype=torch.float32).pin_memory(); hx.uniform_(-1, 1)
hy = torch.empty(N * nvec * 2, dtype=torch.float32).pin_memory()
f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
for _ in range(2): f.work_ptr(hx.data_ptr(), hy.data_ptr(), nvec)
t0 = time.perf_counter()
for _ in range(8): f.work_ptr(hx.data_ptr(), hy.data_ptr(), nvec)
dt = (time.perf_counter() - t0) / 8
print("chunk %s MiB: %.2f ms/step, %.0f Msamples/s, %.1f GB/s each way" % (os.environ.get("CLB200_CHUNK_MB", "8"), dt * 1e3, N * nvec / dt / 1e6, N * nvec * 8 / dt / 1e9))
