"""Per-call latency of work() on scheduler-sized buffers (8192 gr_complex, pageable numpy),
the shape of the reference's own harness (test_clenabled.cc:1237-1251: 1 warm-up + N calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gr_clenabled_b200 import blocks, capi
from oracle import oracle as orc
GPU = (1, 1, 0, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192          # multiple of 8192
x = orc.rng_c32(n, 1); y = orc.rng_c32(n, 2); out = np.zeros(n, np.complex64)
taps = np.hamming(256).astype(np.float32)
cases = {
    "clMultiplyConst": (blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 2.0, capi.OP_MULTIPLY), lambda b: b.work(x, out=out)),
    "clMultiply": (blocks.clMathOp(capi.DTYPE_COMPLEX, *GPU, capi.OP_MULTIPLY), lambda b: b.work(x, y, out=out)),
    "clFFT_8192": (blocks.clFFT(8192, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU), lambda b: b.work(x, out=out)),
    "clFilter_fft_256": (blocks.clFilter(*GPU, 1, taps), lambda b: b.work(x)),
}
for name, (blk, fn) in cases.items():
    fn(blk)
    iters = 200
    t0 = time.perf_counter()
    for _ in range(iters):
        fn(blk)
    dt = (time.perf_counter() - t0) / iters
    print("%-20s %7.1f us/call  %8.1f Msamples/s" % (name, dt * 1e6, n / dt / 1e6), flush=True)
