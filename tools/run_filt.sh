timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_dyn.json 2> gpurun_out/bench_dyn.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_dyn.err
