"""Launches one block's device-resident kernel a few times (for ncu).  usage: prof_one.py fft|fft65536|mathconst|xengine|xengine_batch|xengine_c32|filter|pfb|fir [variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gr_clenabled_b200 import blocks, capi

what = sys.argv[1]
sp = torch.cuda.current_stream().cuda_stream
gpu = (1, 1, 0, 0)
reps = int(os.environ.get("REPS", "3"))
if what in ("fft", "fft4096"):
    N = 8192 if what == "fft" else 4096
    nvec = (1 << 26) // N
    x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *gpu)
    for _ in range(reps):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
elif what == "fft65536":
    N, nvec = 65536, 1024
    x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *gpu)
    for _ in range(reps):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
elif what == "mathconst":
    n = 1 << 26
    x = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *gpu, 0.7071, capi.OP_MULTIPLY)
    for _ in range(reps):
        blk.launch_device(x.data_ptr(), y.data_ptr(), n, sp)
elif what == "xengine":
    A, F, T = 32, 1024, 1024
    npol = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    nb = T * A * F * 2 * npol
    bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(3)]
    vis = torch.empty(F * (A * (A + 1) // 2) * 2 * npol * npol, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(*gpu, False, capi.DTYPE_BYTE, npol, A, 1, 0, F, T, [])
    for i in range(reps):
        blk.launch_device(bufs[i % 3].data_ptr(), vis.data_ptr(), False, sp)
elif what in ("filter", "fir"):
    n = 1 << 26
    x = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    taps = np.hamming(256).astype(np.float32) / 128
    blk = blocks.clFilter(*gpu, 1, taps, 1, 0, what == "fir")
    for _ in range(reps):
        blk.launch_device(x.data_ptr(), n, y.data_ptr(), sp)
elif what == "pfb":
    n = 1 << 26
    x = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    taps = np.hamming(128).astype(np.float32) / 64
    blk = blocks.clPolyphaseChannelizer(*gpu, taps, 65536, 64, 64, list(range(64)))
    for _ in range(reps):
        blk.launch_device(x.data_ptr(), y.data_ptr(), (n - 128) // 64, sp)
elif what == "xengine_batch":
    A, F, T, K = 32, 1024, 1024, 16
    per = T * A * F * 2
    buf = torch.randint(-127, 128, (per * K,), dtype=torch.int8, device="cuda")
    vis = torch.empty(K * F * (A * (A + 1) // 2) * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(*gpu, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    for i in range(reps):
        blk.launch_device_batch(buf.data_ptr(), vis.data_ptr(), K, sp)
elif what == "xengine_packed":
    A, F, T, npol = 16, 1024, 1024, 2
    bufs = [torch.randint(0, 256, (T * A * F * npol,), dtype=torch.uint8, device="cuda") for _ in range(3)]
    vis = torch.empty(F * (A * (A + 1) // 2) * 4 * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(*gpu, False, capi.DTYPE_PACKEDXY, npol, A, 1, 0, F, T, [])
    for i in range(reps):
        blk.launch_device(bufs[i % 3].data_ptr(), vis.data_ptr(), False, sp)
elif what == "xengine_c32":
    A, F, T = 32, 256, 1024
    xc = torch.empty(T * A * F * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    vis = torch.empty(F * (A * (A + 1) // 2) * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(*gpu, False, capi.DTYPE_COMPLEX, 1, A, 1, 0, F, T, [])
    for i in range(reps):
        blk.launch_device(xc.data_ptr(), vis.data_ptr(), False, sp)
torch.cuda.synchronize()
print("done", what)
