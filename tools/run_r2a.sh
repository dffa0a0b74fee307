timeout 120 python -m pytest tests -m gpu -x -q -k "xengine" 2>&1 | tail -3
timeout 200 python - <<'PY'
import sys, json, os
sys.path.insert(0, ".")
import torch, bench
from gr_clenabled_b200 import blocks, capi
for dma in ("0", "1"):
    os.environ["CLB200_XE_INGEST_DMA"] = dma
    r = bench.xengine_host(blocks, capi, 0, torch, None, 1, 1024, "x")
    print("DMA" if dma == "1" else "kernel", {k: (round(v["us_per_integration"]), v["matches_device_launch"]) for k, v in r.items() if isinstance(v, dict)})
    r = bench.xengine_host(blocks, capi, 0, torch, None, 1, 128, "x")
    print("  128ch", {k: (round(v["us_per_integration"]), v["matches_device_launch"]) for k, v in r.items() if isinstance(v, dict)})
PY
