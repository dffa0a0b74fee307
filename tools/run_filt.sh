timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{ echo "clFFT size sweep, device-resident, 64 Mi samples per launch (tools/fft_sweep.py; % of the 6543.7 GB/s copy figure)"; timeout 300 python tools/fft_sweep.py; } > gpurun_out/ev_fft_sweep.txt 2>&1
{ echo "clFFT: static striding vs work-counter tiles (columns form1 = form2 = the per-size default loop form; 16 and 32 points always stride statically), alternating in one process, GB/s (tools/fft_dyn_ab.py)"; timeout 300 python tools/fft_dyn_ab.py 4 5 6 7 8 9 10 11 12 13 14; } > gpurun_out/ev_tile_ab.txt 2>&1
cat gpurun_out/ev_fft_sweep.txt
