for r in 1 0; do for mb in 16 32 64; do CLB200_CHUNK_RAMP=$r CLB200_CHUNK_MB=$mb timeout 100 python tools/e2e_fft.py | sed "s/^/ramp=$r /"; done; done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
