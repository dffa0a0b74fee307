"""clFFT 8192: does a start offset between the two co-resident CTAs of an SM help? (CLB200_FFT_STAGGER_NS)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
N, nvec = 8192, 8192
x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
for ns in (0, 500, 1000, 2000, 3000, 4000, 6000, 0):
    os.environ["CLB200_FFT_STAGGER_NS"] = str(ns)
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
    for _ in range(3): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 / 1e3
    print("stagger %5d ns: %.1f GB/s" % (ns, 16 * N * nvec / t / 1e9), flush=True)
