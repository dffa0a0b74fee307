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
for fc, sl in ((16, 2),):
    os.environ["CLB200_XE_DBG"] = "8"; os.environ["CLB200_XE_SLICES"] = str(sl); os.environ["CLB200_XE_FC"] = str(fc)
    blk = blocks.clXEngine(1, 1, 0, 0, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    for i in range(3): blk.launch_device_i32(bufs[i % 4].data_ptr(), acc.data_ptr(), sp)
    torch.cuda.synchronize()
