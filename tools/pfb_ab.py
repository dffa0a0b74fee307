"""clPolyphaseChannelizer 64 channels x 128 taps, device-resident: runs of consecutive time steps (samples loaded once) vs one step per tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_clenabled_b200 import blocks
sp = torch.cuda.current_stream().cuda_stream
n = 1 << 26
a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
ref = None
for M, ntaps in ((64, 128), (64, 256), (16, 64), (256, 512), (1024, 2048)):
    taps = (np.hamming(ntaps) / ntaps * 2).astype(np.float32)
    niter = (n - ntaps) // M
    for tag, env in (("runs", "1"), ("single", "0"), ("runs", "1"), ("single", "0")):
        os.environ["CLB200_PFB_RUN"] = env
        blk = blocks.clPolyphaseChannelizer(1, 1, 0, 0, taps, 65536, M, M, list(range(M)))
        for _ in range(2): blk.launch_device(a.data_ptr(), b.data_ptr(), niter, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): blk.launch_device(a.data_ptr(), b.data_ptr(), niter, sp)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 / 1e3
        cs = float(b[:1 << 22].double().abs().sum())
        print("M=%4d taps=%4d %-6s %.0f GB/s  checksum %.6f" % (M, ntaps, tag, 16 * niter * M / t / 1e9, cs), flush=True)
