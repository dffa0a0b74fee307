timeout 300 python -m pytest tests -m gpu -x -q -k "xengine" 2>&1 | tail -2
for i in 1 2; do timeout 300 python tools/time_blocks.py 2>&1 | grep -E "complex" | cut -c1-140; done
