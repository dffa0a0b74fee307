"""tools/xe_batch.py -- clXEngine 32 x 1024 x 1024 (BASELINE config 5), device-resident: one launch per integration vs
clb200_xengine_launch_device_batch over K integrations (one grid), CUDA events on the launching stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi

A, F, T = 32, int(os.environ.get("XE_CHANNELS", "1024")), 1024
nbl = A * (A + 1) // 2
per = T * A * F * 2
KMAX = 16
buf = torch.randint(-127, 128, (per * KMAX,), dtype=torch.int8, device="cuda")      # 1 GiB > L2
out = torch.empty(KMAX * F * nbl * 2, dtype=torch.float32, device="cuda")
blk = blocks.clXEngine(1, 2, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
sp = torch.cuda.current_stream().cuda_stream
bytes_per = per + F * nbl * 8


def timed(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def single():
    for k in range(KMAX):
        blk.launch_device(buf.data_ptr() + k * per, out.data_ptr() + k * F * nbl * 8, False, sp)

us = timed(single, 10) / KMAX
print("one launch per integration (PDL)  %6.2f us / integration  %5.0f GB/s" % (us, bytes_per / us / 1e3))
for K in (2, 4, 8, 16):
    us = timed(lambda: blk.launch_device_batch(buf.data_ptr(), out.data_ptr(), K, sp), 10) / K
    print("batch of %2d in one grid            %6.2f us / integration  %5.0f GB/s" % (K, us, bytes_per / us / 1e3))
