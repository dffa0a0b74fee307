"""clFFT 8192 A/B of build-time-equal kernel variants selected by env at block creation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_clenabled_b200 import blocks, capi
from oracle import oracle as orc
sp = torch.cuda.current_stream().cuda_stream
N, nvec = 8192, 8192
x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
xs = orc.rng_c32(N * 4, 7)
want = orc.fft(xs, N, -1)
for tag, env in (("twiddles from L1 (default)", {}), ("pass-1 twiddles in smem", {"CLB200_FFT_TW1S": "1"}),
                 ("twiddles from L1 (default)", {}), ("pass-1 twiddles in smem", {"CLB200_FFT_TW1S": "1"})):
    for k in ("CLB200_FFT_TW1S",): os.environ.pop(k, None)
    os.environ.update(env)
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
    got = f.work(xs)
    err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    for _ in range(3): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 / 1e3
    print("%-30s %.1f GB/s   rel err %.2e" % (tag, 16 * N * nvec / t / 1e9, err), flush=True)
