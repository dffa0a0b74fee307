timeout 300 python -m pytest tests -m gpu -x -q -k "secondary or elementwise or mathconst or mathop or log or mag" 2>&1 | tail -3
timeout 100 python - <<'PY'
import sys
sys.path.insert(0, ".")
import torch, bench
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
n = 1 << 26
a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(0.1, 1)
b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
lib = capi.load()
for name, blk, nit, by in (("clComplexToMag", blocks.clComplexToMag(1, 2, 0, 0), n, 12), ("clComplexToArg", blocks.clComplexToArg(1, 2, 0, 0), n, 12), ("clLog10", blocks.clLog(1, 2, 0, 0, 10.0, 0.0), 2 * n, 8)):
    t = bench._timeit(torch, lambda: capi.check(lib.clb200_unary_launch_device(blk._h, a.data_ptr(), b.data_ptr(), nit, sp)))
    print("%-16s %6.0f GB/s  %.1f%% of HBM" % (name, by * nit / t / 1e9, by * nit / t / 1e9 / 65.49))
PY
