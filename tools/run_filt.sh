for c in 2 3 4 5 6 8; do echo "ctas/SM target $c"; CLB200_XE_C32_CTAS=$c timeout 300 python tools/time_blocks.py 2>&1 | grep -E "complex" | cut -c1-120; done
