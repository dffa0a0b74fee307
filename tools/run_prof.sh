# round-2 profile evidence: launch list of the bench command + one full capture per kernel (ncu replays each kernel
# ~40 times: one GPU, never under torchrun)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
cap() { REPS=3 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${4:-1} -c 1 -f -o gpurun_out/$2 python tools/prof_one.py $3 > gpurun_out/ncu_$2.log 2>&1; echo "$2 rc=$?"; }
cap k_xengine_tma xe_tma xengine
cap k_xengine_tma xe_batch xengine_batch
cap k_xengine_tma xe_pk xengine_packed
cap k_xengine_c32 xe_c32 xengine_c32
cap "^k_fft$" fft fft
cap "^k_fft$" fft4096 fft4096
cap k_fftfilt fftfilt filter
cap k_fir fir fir
cap k_pfb pfb pfb
cap k_map1 map1 mathconst
cap k_fft_col fftcolA fft65536 2     # launches alternate pass A, pass B: the third is pass A of the second repetition
cap k_fft_col fftcolB fft65536 3
