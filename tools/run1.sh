mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k xengine > gpurun_out/t_xe.log 2>&1; echo "xe tests rc=$?"
tail -5 gpurun_out/t_xe.log
timeout 300 python tools/xe_tune.py > gpurun_out/xe_tune.log 2>&1; cat gpurun_out/xe_tune.log | tail -20
timeout 300 python tools/xe_dbg2.py > gpurun_out/xe_dbg2.log 2>&1; cat gpurun_out/xe_dbg2.log | tail -20
