"""FFT-filter A/B (256 taps, 64 Mi samples device-resident): register cap / block-size variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_clenabled_b200 import blocks, capi
from oracle import oracle as orc
sp = torch.cuda.current_stream().cuda_stream
n = 1 << 26
a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
taps = np.zeros(256, np.float32); taps[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
ref = None
for tag, env in (("168 regs, 12 warps/SM", {}), ("128 regs, 16 warps/SM", {"CLB200_FILT_MINB": "16"}), ("NF=4096", {"CLB200_FILT_NF": "4096"}),
                 ("168 regs, 12 warps/SM", {}), ("128 regs, 16 warps/SM", {"CLB200_FILT_MINB": "16"})):
    for k in ("CLB200_FILT_MINB", "CLB200_FILT_NF"): os.environ.pop(k, None)
    os.environ.update(env)
    blk = blocks.clFilter(1, 1, 0, 0, 1, taps, 1, 0, False)
    for _ in range(2): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 / 1e3
    print("%-26s %.1f GB/s  (%.1f Gsamples/s)" % (tag, 16 * n / t / 1e9, n / t / 1e9), flush=True)
