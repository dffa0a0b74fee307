timeout 600 python -m pytest tests -m gpu -x -q -k "fft or FFT or zero or bad or ctor or argument" 2>&1 | tail -8
