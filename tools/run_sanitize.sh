# compute-sanitizer passes over a subset of the GPU parity tests (the full suite under the sanitizer takes too long):
# memcheck on every block's small-shape tests, racecheck + synccheck on the X-engine kernels (cluster / DSMEM exchange,
# mbarrier rings) and the FFT / filter kernels (shared-memory exchanges).
mkdir -p gpurun_out
SEL='fft_forward_all_sizes or fft_real_input_sizes or filter_tap_counts or pfb_vs_oracle or xengine_ichar_bit_exact or xengine_tma_feed_ragged or xengine_streaming_push_poll or xengine_batched or xengine_complex_float or mathconst or mathop'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_pytest.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_pytest.txt; tail -4 gpurun_out/sanitizer_memcheck.log
SEL2='xengine_ichar_bit_exact or xengine_tma_feed_ragged or xengine_batched or fft_backward_window_shift or filter_lowpass_256 or xengine_complex_float or pfb_vs_oracle'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL2" > gpurun_out/sanitizer_racecheck_pytest.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_pytest.txt; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/sanitizer_synccheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL2" > gpurun_out/sanitizer_synccheck_pytest.txt 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/sanitizer_synccheck_pytest.txt; tail -4 gpurun_out/sanitizer_synccheck.log
