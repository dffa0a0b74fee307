import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
A, F, T = 32, 1024, 1024
nb = T * A * F * 2
bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(4)]
nout = F * (A * (A + 1) // 2)
acc = torch.empty(nout * 2, dtype=torch.int32, device="cuda")
def run(tag, **env):
    for k in list(os.environ):
        if k.startswith("CLB200_XE_"): os.environ.pop(k)
    for k, v in env.items(): os.environ["CLB200_XE_" + k] = str(v)
    blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    for i in range(3): blk.launch_device_i32(bufs[i % 4].data_ptr(), acc.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16): blk.launch_device_i32(bufs[i % 4].data_ptr(), acc.data_ptr(), sp)
    e1.record(); torch.cuda.synchronize()
    print("%-40s %7.1f us" % (tag, e0.elapsed_time(e1) / 16 * 1e3), flush=True)
run("fc=16 sl=2 full", FC=16, SLICES=2)
run("fc=16 sl=2 no epilogue", FC=16, SLICES=2, DBG=4)
run("fc=8 full", FC=8)
run("default", )
os.environ["CLB200_XE_DBG"] = "8"; os.environ["CLB200_XE_SLICES"] = "2"; os.environ["CLB200_XE_FC"] = "16"
blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
for i in range(3): blk.launch_device_i32(bufs[i % 4].data_ptr(), acc.data_ptr(), sp)
torch.cuda.synchronize()
