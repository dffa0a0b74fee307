timeout 300 python -m pytest tests -m gpu -x -q -k "not_a_power or multi_kernel" 2>&1 | tail -3
CLB200_FFT_CZ_UNFUSED=1 timeout 300 python -m pytest tests -m gpu -x -q -k "not_a_power" 2>&1 | tail -1
timeout 300 python tools/fft_cz_ab.py
