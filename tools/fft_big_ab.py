"""clFFT above 16384 points, device-resident: the two-pass column kernels vs the five-pass four-step path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
total = 1 << 26
x = torch.empty(total * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
for logn in range(15, 21):
    N = 1 << logn
    row = []
    for four in ("0", "1"):
        os.environ["CLB200_FFT_FOURSTEP"] = four
        f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
        for _ in range(2): f.launch_device(x.data_ptr(), y.data_ptr(), total // N, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f.launch_device(x.data_ptr(), y.data_ptr(), total // N, sp)
        e1.record(); torch.cuda.synchronize()
        row.append(16 * total / (e0.elapsed_time(e1) / 5) / 1e6)
    print("N=%8d  two-pass (where built) %6.0f GB/s   four-step %6.0f GB/s" % (N, row[0], row[1]), flush=True)
