timeout 200 python -m pytest tests -m gpu -x -q -k "xengine" 2>&1 | tail -8
timeout 100 python - <<'PY'
import sys, os
sys.path.insert(0, ".")
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
A, F, T, npol = 16, 1024, 1024, 2
nb = T * A * F * npol
for env in ("0", "1"):
    os.environ["CLB200_XE_UNPACK_PASS"] = env
    bufs = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device="cuda") for _ in range(8)]
    vis = torch.empty(F * (A * (A + 1) // 2) * 4 * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(1, 2, 0, 0, False, capi.DTYPE_PACKEDXY, npol, A, 1, 0, F, T, [])
    for i in range(3): blk.launch_device(bufs[i].data_ptr(), vis.data_ptr(), False, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(32): blk.launch_device(bufs[i % 8].data_ptr(), vis.data_ptr(), False, sp)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 32 * 1e3
    print("packed 16st x 2pol x 1024ch x 1024t, %s: %.1f us / integration" % ("separate unpack pass" if env == "1" else "fused", us))
PY
