set -x
timeout 600 python -m pytest tests/test_gpu_sectors.py -x -q > gpurun_out/pytest_sectors4.log 2>&1; tail -4 gpurun_out/pytest_sectors4.log
rm -f gpurun_out/sector_r01.jsonl
timeout 300 python scripts/sector_bench.py hub34_half hub44_10 hub44_half mol24 mol28 2>&1 | grep -o '"case[^}]*ms_per_apply": [0-9.]*'
cp gpurun_out/sector_r01.jsonl gpurun_out/sector_r01_v4.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sector_apply -s 4 -c 1 -o gpurun_out/k_sector_r01v4 -f python scripts/sector_bench.py hub44_half > gpurun_out/ncu_sector.log 2>&1; tail -3 gpurun_out/ncu_sector.log
