timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches','vs_baseline','dtype')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
