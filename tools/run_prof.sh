# round-1 profile evidence: launch list of the bench command + one full capture per headline kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_xengine_tma -s 1 -c 1 -f -o gpurun_out/xe_tma python tools/prof_one.py xengine > gpurun_out/ncu_xe.log 2>&1; echo "xe rc=$?"
if [ -n "$PROF_ALL" ]; then
REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fft -s 1 -c 1 -f -o gpurun_out/fft python tools/prof_one.py fft > gpurun_out/ncu_fft.log 2>&1; echo "fft rc=$?"
REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fftfilt -s 1 -c 1 -f -o gpurun_out/fftfilt python tools/prof_one.py filter > gpurun_out/ncu_filt.log 2>&1; echo "filt rc=$?"
REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pfb -s 1 -c 1 -f -o gpurun_out/pfb python tools/prof_one.py pfb > gpurun_out/ncu_pfb.log 2>&1; echo "pfb rc=$?"
fi
