for pk in 1 0 1 0; do CLB200_FIR_PACKED=$pk timeout 100 python tools/fir_ab.py 2>&1 | tail -1 | sed "s/^/packed=$pk /"; done
timeout 300 python -m pytest tests -m gpu -x -q -k "filter or Filter or fir or dynamic" 2>&1 | tail -3
