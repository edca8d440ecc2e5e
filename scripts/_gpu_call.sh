set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01d.log 2>&1; tail -3 gpurun_out/pytest_gpu_r01d.log
timeout 300 python scripts/quick_apply_bench.py mol24 mol28 hub30 2>&1 | tail -3 | cut -c1-90
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply8 -s 2 -c 1 -o gpurun_out/k_apply_r01v7_n24 -f python scripts/quick_apply_bench.py mol24 > gpurun_out/ncu_v7.log 2>&1; tail -2 gpurun_out/ncu_v7.log
