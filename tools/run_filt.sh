timeout 300 python -m pytest tests -m gpu -x -q -k "above_16384 or multi_kernel or not_a_power" 2>&1 | tail -2
timeout 300 python tools/fft_big_ab.py 2>&1
