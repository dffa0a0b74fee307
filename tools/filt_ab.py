"""FFT-filter A/B (256 taps, 64 Mi samples device-resident): kernel variants selected by environment switches.
Every variant's output is compared with the first one's (max abs difference)."""
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
KEYS = ("CLB200_FILT_MINB", "CLB200_FILT_NF", "CLB200_FILT_PF", "CLB200_FILT_COMPACT")
variants = [("base", {})]
for spec in sys.argv[1:]:
    variants.append((spec, dict(kv.split("=") for kv in spec.split(","))))
variants.append(("base again", {}))
ref = None
for tag, env in variants:
    for k in KEYS: os.environ.pop(k, None)
    os.environ.pop("CLB200_STATIC_TILES", None)
    os.environ.update({("CLB200_STATIC_TILES" if k == "STATIC" else "CLB200_FILT_" + k): v for k, v in env.items()})
    blk = blocks.clFilter(1, 1, 0, 0, 1, taps, 1, 0, False)
    for _ in range(2): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): blk.launch_device(a.data_ptr(), n, b.data_ptr(), sp)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 / 1e3
    # a fresh block (zero history) for the comparison run
    blk2 = blocks.clFilter(1, 1, 0, 0, 1, taps, 1, 0, False)
    b.zero_(); blk2.launch_device(a.data_ptr(), n, b.data_ptr(), sp); torch.cuda.synchronize()
    if ref is None: ref = b.clone(); d = 0.0
    else: d = float((b - ref).abs().max())
    print("%-28s %.1f GB/s  (%.1f Gsamples/s)  maxdiff vs base %.2e" % (tag, 16 * n / t / 1e9, n / t / 1e9, d), flush=True)
