set -x
timeout 600 python -m pytest tests/test_gpu_sectors.py -x -q > gpurun_out/pytest_sectors2.log 2>&1; tail -4 gpurun_out/pytest_sectors2.log
timeout 300 python scripts/sector_bench.py hub34_half hub44_10 hub44_half mol24 mol28 2>&1 | tail -6
