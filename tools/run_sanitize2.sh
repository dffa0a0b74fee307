# compute-sanitizer over the kernels added in the second half of round 2: work-counter tiles (hand-over through shared
# memory), the column-transform FFT passes, the chirp-z path, the rewritten FIR loop
mkdir -p gpurun_out
SEL='dynamic_tiles or above_16384_four_step[32768] or above_16384_four_step[65536] or not_a_power_of_two[12] or not_a_power_of_two[1000] or not_a_power_of_two[10000]'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer2_$tool.log python -m pytest tests/test_dynamic_tiles.py tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer2_${tool}_pytest.txt 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer2_${tool}_pytest.txt; tail -3 gpurun_out/sanitizer2_$tool.log
done
