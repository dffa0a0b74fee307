"""clFFT lengths that are not a power of two, device-resident: fused two-kernel chirp-z vs the five-kernel form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
total = 1 << 25
x = torch.empty(total * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
for N in (12, 100, 1000, 1536, 3000, 6000, 8000, 10000):
    row = []
    for unf in ("0", "1"):
        os.environ["CLB200_FFT_CZ_UNFUSED"] = unf
        f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
        nv = total // N
        for _ in range(2): f.launch_device(x.data_ptr(), y.data_ptr(), nv, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f.launch_device(x.data_ptr(), y.data_ptr(), nv, sp)
        e1.record(); torch.cuda.synchronize()
        row.append(16 * nv * N / (e0.elapsed_time(e1) / 5) / 1e6)
    print("N=%6d  fused %6.0f GB/s   five kernels %6.0f GB/s" % (N, row[0], row[1]), flush=True)
