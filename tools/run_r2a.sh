timeout 200 python -m pytest tests -m gpu -x -q -k "xengine_complex" 2>&1 | tail -2
timeout 100 python - <<'PY'
import sys, os
sys.path.insert(0, ".")
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
T = 1024
for F, npol, AA in ((256, 1, 32), (1024, 1, 32), (256, 2, 16), (64, 1, 32)):
    xc = torch.empty(T * AA * F * npol * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    nb = AA * (AA + 1) // 2
    vis = torch.empty(F * nb * npol * npol * 2, dtype=torch.float32, device="cuda")
    blk = blocks.clXEngine(1, 2, 0, 0, False, capi.DTYPE_COMPLEX, npol, AA, 1, 0, F, T, [])
    for _ in range(3): blk.launch_device(xc.data_ptr(), vis.data_ptr(), False, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): blk.launch_device(xc.data_ptr(), vis.data_ptr(), False, sp)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    by = T * AA * F * npol * 8 + vis.numel() * 4
    print("complex F=%d npol=%d A=%d: %.1f us  %.0f GB/s  %.1f TFLOP/s" % (F, npol, AA, us, by / us / 1e3, 8.0 * nb * npol * npol * T * F / us / 1e6))
PY
