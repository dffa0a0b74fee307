"""X-engine time vs integration length (slope = streaming rate, intercept = fixed launch/prologue/epilogue cost)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
A, F = 32, 1024
Tmax = 4096
nbmax = Tmax * A * F * 2
bufs = [torch.randint(-127, 128, (nbmax,), dtype=torch.int8, device="cuda") for _ in range(2)]
nout = F * (A * (A + 1) // 2)
acc = torch.empty(nout * 2, dtype=torch.int32, device="cuda")
def run(tag, T, **env):
    for k in list(os.environ):
        if k.startswith("CLB200_XE_"): os.environ.pop(k)
    for k, v in env.items(): os.environ["CLB200_XE_" + k] = str(v)
    blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    for i in range(3): blk.launch_device_i32(bufs[i % 2].data_ptr(), acc.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16): blk.launch_device_i32(bufs[i % 2].data_ptr(), acc.data_ptr(), sp)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 16 * 1e3
    print("%-34s T=%5d %7.1f us  %6.0f GB/s" % (tag, T, us, T * A * F * 2 / us / 1e3), flush=True)
for tag, env in (("default (16 ch/CTA, clusters of 2 slices)", dict()), ("8 ch/CTA, no slicing", dict(FC=8)),
                 ("no programmatic dependent launch", dict(PDL=0)), ("ldg-fed kernel (r1)", dict(TMA=0))):
    for T in (256, 512, 1024, 2048, 4096):
        run(tag, T, **env)
