timeout 100 python tools/fir_ab.py 2>&1 | tail -1
CLB200_FIR_CTAS=4 timeout 100 python tools/fir_ab.py 2>&1 | tail -1
CLB200_FIR_CTAS=6 timeout 100 python tools/fir_ab.py 2>&1 | tail -1
CLB200_FIR_CTAS=8 timeout 100 python tools/fir_ab.py 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q -k "filter or Filter or fir or dynamic" 2>&1 | tail -3
