mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k xengine > gpurun_out/t_xe.log 2>&1; echo "xe tests rc=$?"; tail -3 gpurun_out/t_xe.log
timeout 300 python tools/xe_tune.py 2>&1 | tee gpurun_out/xe_tune.log | tail -12
timeout 300 python tools/xe_scale.py 2>&1 | tee gpurun_out/xe_scale.log | tail -24
