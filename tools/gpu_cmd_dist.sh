timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tree_parallel" 2>&1 | tail -15
