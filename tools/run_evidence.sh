# measurement evidence for DESIGN.md §4.0 / §4.1 / §4.2 / §4.3 (text files copied into profiles/)
mkdir -p gpurun_out
{ echo "clFFT size sweep, device-resident, 64 Mi samples per launch (tools/fft_sweep.py; % of the 6543.7 GB/s copy figure)"; timeout 300 python tools/fft_sweep.py; } > gpurun_out/ev_fft_sweep.txt 2>&1
{ echo "clFFT: static striding vs the two work-counter loop forms, alternating in one process, GB/s (tools/fft_dyn_ab.py)"; timeout 300 python tools/fft_dyn_ab.py 7 8 9 10 11 12 13 14; } > gpurun_out/ev_tile_ab.txt 2>&1
{ echo "clFFT above 16384 points: two-pass column kernels vs the five-pass four-step path (tools/fft_big_ab.py)"; timeout 300 python tools/fft_big_ab.py; } > gpurun_out/ev_fft_big.txt 2>&1
{ echo "256-tap FFT filter, 64 Mi samples: kernel variants by switch (tools/filt_ab.py)"; timeout 300 python tools/filt_ab.py STATIC=1 PF=1 PF=2 MINB=13 MINB=14 MINB=15 MINB=16 COMPACT=1 COMPACT=2 NF=4096; } > gpurun_out/ev_filt_ab.txt 2>&1
{ echo "register-only butterfly probe (tools/src/bfly_probe.cu)"; tools/bin/bfly_probe; } > gpurun_out/ev_bfly_probe.txt 2>&1
{ echo "polyphase channelizer: runs of consecutive time steps vs one step per tile (tools/pfb_ab.py)"; timeout 300 python tools/pfb_ab.py; } > gpurun_out/ev_pfb_ab.txt 2>&1
{ echo "time-domain FIR, 256 taps: packed FFMA2 vs scalar FFMA (tools/fir_ab.py)"; for pk in 1 0; do CLB200_FIR_PACKED=$pk timeout 100 python tools/fir_ab.py | sed "s/^/CLB200_FIR_PACKED=$pk /"; done; } > gpurun_out/ev_fir_ab.txt 2>&1
tail -n 3 gpurun_out/ev_*.txt
