python -m pytest tests -m gpu -x -q -k "xengine" 2>&1 | tail -3
for d in 0 4; do echo "== DBG=$d"; CLB200_XE_DBG=$d python tools/xe_batch.py 2>&1 | tail -5; done
