for s in 0 1; do echo "== STATIC=$s"; CLB200_STATIC_TILES=$s timeout 300 python tools/time_blocks.py 2>&1 | grep -E "^(clMultiply|clComplex|clLog)" | cut -c1-200; done
timeout 600 python -m pytest tests -m gpu -x -q -k "mathconst or mathop or log or snr or mag or arg or unary or secondary or dynamic or zero" 2>&1 | tail -3
