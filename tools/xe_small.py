"""X-engine decomposition sweep for the per-rank shapes of the channel-sharded runs (32 stations, 1024 time steps,
F = 512 / 256 / 128 channels): channels per CTA x time slices, CUDA-event time per launch (PDL, 32 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi

sp = torch.cuda.current_stream().cuda_stream
A, T = 32, 1024
for F in (1024, 512, 256, 128, 64, 2048):
    nb = T * A * F * 2
    bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(8)]
    nout = F * (A * (A + 1) // 2)
    vis = torch.empty(nout * 2, dtype=torch.float32, device="cuda")
    ref = None
    for fc, sl in (((0, None),) if os.environ.get('XE_SMALL_DEFAULT_ONLY') else ((0, None), (16, 1), (16, 2), (16, 4), (16, 8), (8, 1), (8, 2), (8, 4), (8, 8))):
        for k in ("CLB200_XE_SLICES", "CLB200_XE_FC"):
            os.environ.pop(k, None)
        if fc:
            os.environ["CLB200_XE_FC"] = str(fc)
            os.environ["CLB200_XE_SLICES"] = str(sl)
        blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
        blk.launch_device(bufs[0].data_ptr(), vis.data_ptr(), False, sp)
        torch.cuda.synchronize()
        if ref is None:
            ref = vis.clone()
        ok = bool(torch.equal(vis, ref))
        for i in range(4):
            blk.launch_device(bufs[i % 8].data_ptr(), vis.data_ptr(), False, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(32):
            blk.launch_device(bufs[i % 8].data_ptr(), vis.data_ptr(), False, sp)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 32 * 1e3
        print("F=%4d fc=%-2s slices=%-4s %6.1f us  same=%s" % (F, fc or "dflt", sl, us, ok), flush=True)
