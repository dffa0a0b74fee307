timeout 300 python -m pytest tests -m gpu -x -q -k "fft" 2>&1 | tail -5
timeout 100 python - <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from gr_clenabled_b200 import blocks, capi
sp = torch.cuda.current_stream().cuda_stream
for N in (16384, 32768, 65536, 1 << 20):
    nvec = (1 << 26) // N
    x = torch.empty(N * nvec * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    f = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 2, 0, 0)
    for _ in range(2): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 / 1e3
    print("clFFT %8d-pt: %7.1f Gsamples/s  %5.0f GB/s" % (N, N * nvec / t / 1e9, 16 * N * nvec / t / 1e9))
PY
