"""Device-resident throughput of clFFT for every supported size (512 MiB batches, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
total = 1 << 26
x = torch.empty(total * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
for logn in range(1, 15):
    N = 1 << logn
    nvec = total // N
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
    for _ in range(2):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("N=%5d  %7.1f Gsamples/s  %6.0f GB/s  %4.1f%% of HBM" % (N, total / ms / 1e6, 16 * total / ms / 1e6, 16 * total / ms / 1e6 / 65.437), flush=True)
