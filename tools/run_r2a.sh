mkdir -p gpurun_out
SEL2='xengine_ichar_bit_exact or xengine_tma_feed_ragged or xengine_batched or fft_backward_window_shift or filter_lowpass_256 or xengine_complex_float'
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/sanitizer_synccheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL2" > gpurun_out/sanitizer_synccheck_pytest.txt 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/sanitizer_synccheck_pytest.txt; tail -4 gpurun_out/sanitizer_synccheck.log
