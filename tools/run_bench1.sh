mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
b = l.pop("blocks", {})
print(json.dumps(l)[:1800])
for k, v in b.items():
    print(k, json.dumps(v)[:900])
PY
