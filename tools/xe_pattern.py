"""X-engine access-pattern experiment: same bytes, different channel counts (row length)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
A = 32
for F, T in [(1024, 1024), (256, 4096), (64, 16384), (16, 65536)]:
    nb = T * A * F * 2
    bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(4)]
    vis = torch.empty(F * (A * (A + 1) // 2) * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    for i in range(3):
        blk.launch_device(bufs[i % 4].data_ptr(), vis.data_ptr(), False, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        blk.launch_device(bufs[i % 4].data_ptr(), vis.data_ptr(), False, sp)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 8 * 1e3
    print("F=%d T=%d: %.1f us, %.0f GB/s" % (F, T, us, nb / us / 1e3), flush=True)
    del bufs
