"""clFFT device-resident: static striding vs the two work-counter loop forms, interleaved in one process (same buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
if os.environ.get("CLB200_LIB"): capi.LIB_PATH = os.environ["CLB200_LIB"]
total = 1 << 26
x = torch.empty(total * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
def timed(f, nvec, reps=10):
    for _ in range(2): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record(); torch.cuda.synchronize()
    return 16 * total / (e0.elapsed_time(e1) / reps) / 1e6
for logn in [int(a) for a in sys.argv[1:]] or range(4, 15):
    N = 1 << logn
    fs = {}
    for form in ("1", "2"):
        os.environ["CLB200_FFT_LOOP"] = form
        fs[form] = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
    res = {"static": [], "form1": [], "form2": []}
    for rep in range(3):
        os.environ["CLB200_STATIC_TILES"] = "1"
        res["static"].append(timed(fs["1"], total // N))
        os.environ["CLB200_STATIC_TILES"] = "0"
        res["form1"].append(timed(fs["1"], total // N))
        res["form2"].append(timed(fs["2"], total // N))
    print("N=%5d  " % N + "   ".join("%s %s" % (k, " ".join("%5.0f" % v for v in r)) for k, r in res.items()), flush=True)
