N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
b = l.pop("blocks", {})
print(json.dumps(l)[:1500])
for k, v in b.items():
    print(k, json.dumps(v)[:1200])
PY
