"""X-engine feed/decomposition sweep at the BASELINE config (32 stations x 1024 channels x 1024 steps, IChar):
CUDA-event time per integration for the LDG-fed and TMA-fed tcgen05 kernels, time-slice counts and TMA L2
promotion.  Env knobs are read when the block is created.  usage: xe_tune.py [npol]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi

sp = torch.cuda.current_stream().cuda_stream
npol = int(sys.argv[1]) if len(sys.argv) > 1 else 1
A, F, T = 32 // npol, 1024, 1024
nb = T * A * F * npol * 2
bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(4)]
nout = F * (A * (A + 1) // 2) * npol * npol
vis = torch.empty(nout * 2, dtype=torch.float32, device="cuda")
acc = torch.empty(nout * 2, dtype=torch.int32, device="cuda")
ref = None


def run(tag, **env):
    global ref
    for k in ("CLB200_XE_TMA", "CLB200_XE_SLICES", "CLB200_XE_L2PROMO", "CLB200_XE_LEGACY", "CLB200_XE_FC", "CLB200_XE_PDL"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["CLB200_XE_" + k] = str(v)
    blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, npol, A, 1, 0, F, T, [])
    blk.launch_device_i32(bufs[0].data_ptr(), acc.data_ptr(), sp)
    torch.cuda.synchronize()
    if ref is None:
        ref = acc.clone()
    ok = bool(torch.equal(acc, ref))
    for i in range(3):
        blk.launch_device(bufs[i % 4].data_ptr(), vis.data_ptr(), False, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16):
        blk.launch_device(bufs[i % 4].data_ptr(), vis.data_ptr(), False, sp)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 16 * 1e3
    print("%-28s %7.1f us  %6.0f GB/s  same_as_first=%s" % (tag, us, (nb + nout * 8) / us / 1e3, ok), flush=True)


run("ldg feed (r1 kernel)", TMA=0)
run("tma default")
run("tma default, no PDL", PDL=0)
run("tma fc=8", FC=8)
run("tma fc=8, no PDL", FC=8, PDL=0)
run("tma fc=16 slices=2", FC=16, SLICES=2)
run("tma fc=16 slices=2, no PDL", FC=16, SLICES=2, PDL=0)
run("tma default (again)")
