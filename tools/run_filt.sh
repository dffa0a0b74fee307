timeout 300 python -m pytest tests -m gpu -x -q -k "above_16384 or not_a_power" 2>&1 | tail -4
timeout 300 python tools/fft_big_ab.py 2>&1 | head -6
