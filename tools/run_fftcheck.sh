timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_next_rows.py -m gpu -q -k "fft or filter or pfb or polyphase or xcorr" 2>&1 | tail -2
python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fft', round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
for k,v in d['blocks'].items():
    if 'frac_hbm' in v: print(k, round(v['GBps']), round(v['frac_hbm'],3))
"
