N=${1:-2}
bash tools/run_bench2.sh $N
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400 | tee gpurun_out/bench_ref_n$N.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe.py 2>&1 | tail -1 | tee gpurun_out/h2d_probe_n$N.json
