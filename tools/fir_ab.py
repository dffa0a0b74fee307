"""Time-domain FIR (256 taps, 64 Mi samples device-resident): resident CTAs per SM (CLB200_FIR_CTAS, read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_clenabled_b200 import blocks
sp = torch.cuda.current_stream().cuda_stream
n = 1 << 26
a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
taps = (np.hamming(256) / 128).astype(np.float32)
blk = blocks.clFilter(1, 1, 0, 0, 1, taps, 1, 0, True)
for _ in range(2): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3 / 1e3
print("CLB200_FIR_CTAS=%s  %.1f Gsamples/s  %.1f TFLOP/s fp32  checksum %.6f" % (os.environ.get("CLB200_FIR_CTAS", "default"), n / t / 1e9, n * 1024 / t / 1e12, float(b[:1 << 20].double().sum())), flush=True)
