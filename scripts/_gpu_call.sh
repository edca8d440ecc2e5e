timeout 300 python scripts/quick_apply_bench.py mol24 mol26 hub28 rand24c 2>&1 | tail -4 | cut -c1-90
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
