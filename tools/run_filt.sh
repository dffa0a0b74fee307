timeout 600 python -m pytest tests -m gpu -x -q -k "multi_kernel_sizes" 2>&1 | tail -5
