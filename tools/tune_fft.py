"""Times the alternative instantiations of the 8192-pt FFT kernel (CLB200_FFT_VARIANT)
device-resident, CUDA events, inputs larger than L2.  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gr_clenabled_b200 import blocks, capi

N, nvec = 8192, 8192
x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
y = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
res = {}
names = {0: "EPT32 256thr minb2 (default)", 1: "EPT16 512thr minb2", 2: "EPT32 256thr minb1", 3: "EPT16 512thr minb1",
         4: "EPT8 1024thr minb1", 5: "EPT8 1024thr minb2"}
for var in [int(a) for a in sys.argv[1:]] or sorted(names):
    os.environ["CLB200_FFT_VARIANT"] = str(var)
    try:
        f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 1, 0, 0)
    except Exception as e:
        res[var] = str(e)
        continue
    for _ in range(3):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = 16 * N * nvec / ms / 1e6
    res[var] = {"name": names.get(var, "?"), "ms": ms, "GBps": gbs, "Gsamples_s": N * nvec / ms / 1e6}
    print(var, res[var], flush=True)
print(json.dumps(res))
