set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01e.log 2>&1; tail -3 gpurun_out/pytest_gpu_r01e.log
timeout 600 python bench.py > gpurun_out/bench_n1_r01c.json 2> gpurun_out/bench_n1_r01c.err; tail -c 400 gpurun_out/bench_n1_r01c.json
timeout 600 python scripts/sweep.py > gpurun_out/sweep_r01b.log 2>&1; tail -3 gpurun_out/sweep_r01b.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_n28_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ncu_launches_v7.log 2>&1; tail -2 gpurun_out/ncu_launches_v7.log | cut -c1-300
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_apply_tile -s 3 -c 1 --csv --log-file gpurun_out/k_apply_r01v7_n28_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ncu_traffic_v7.log 2>&1; tail -3 gpurun_out/k_apply_r01v7_n28_traffic.csv
