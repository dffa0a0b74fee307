echo "== flag sync"; timeout 100 python tools/percall.py 2>&1 | tail -4
echo "== cudaStreamSynchronize"; CLB200_FLAG_SYNC=0 timeout 100 python tools/percall.py 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
