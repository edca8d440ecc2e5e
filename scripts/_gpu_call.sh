set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01c.log 2>&1; tail -4 gpurun_out/pytest_gpu_r01c.log
rm -f gpurun_out/sector_r01.jsonl
timeout 300 python scripts/sector_bench.py hub34_half hub44_10 hub44_half --lanczos 2>&1 | grep -o '"case[^}]*'
timeout 200 python scripts/sector_bench.py mol24 mol28 2>&1 | grep -o '"case[^}]*ms_per_apply": [0-9.]*'
cp gpurun_out/sector_r01.jsonl gpurun_out/sector_r01_final.jsonl
timeout 600 python bench.py > gpurun_out/bench_n1_r01b.json 2> gpurun_out/bench_n1_r01b.err; tail -c 1500 gpurun_out/bench_n1_r01b.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_r01b.json 2>&1; tail -c 600 gpurun_out/bench_ref_r01b.json
