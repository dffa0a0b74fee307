timeout 300 python -m pytest tests -m gpu -x -q -k "xengine" 2>&1 | tail -3
XE_SMALL_DEFAULT_ONLY=1 timeout 100 python tools/xe_small.py
echo "== with the wait"; CLB200_XE_PDL_WAIT=1 XE_SMALL_DEFAULT_ONLY=1 timeout 100 python tools/xe_small.py
