mkdir -p gpurun_out
SEL='not_a_power_of_two[12] or not_a_power_of_two[1000] or not_a_power_of_two[3125] or pfb_runs_of_time_steps or pfb_vs_oracle or xengine_complex_float or secondary or unary or log or mag'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer3_$tool.log python -m pytest tests/test_gpu_parity.py tests/test_next_rows.py -x -q -k "$SEL" > gpurun_out/sanitizer3_${tool}_pytest.txt 2>&1; echo "$tool rc=$?"; tail -1 gpurun_out/sanitizer3_${tool}_pytest.txt; tail -1 gpurun_out/sanitizer3_$tool.log
done
cap() { REPS=3 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${4:-1} -c 1 -f -o gpurun_out/$2 python tools/prof_one.py $3 > gpurun_out/ncu_$2.log 2>&1; echo "$2 rc=$?"; }
cap k_xengine_c32 xe_c32 xengine_c32
