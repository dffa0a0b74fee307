N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 120 $TR --master-port 29512 tools/h2d_probe.py 2>&1 | tail -1 | tee gpurun_out/h2d_probe_n$N.json
timeout 400 $TR --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
b = l.pop("blocks", {})
print(json.dumps(l["e2e"]), l["value"])
for k, v in b.items():
    print(k, json.dumps(v)[:700])
PY
