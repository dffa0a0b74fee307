cap() { REPS=3 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${4:-1} -c 1 -f -o gpurun_out/$2 python tools/prof_one.py $3 > gpurun_out/ncu_$2.log 2>&1; echo "$2 rc=$?"; }
cap "^k_fft$" fft fft
ls -la gpurun_out/
