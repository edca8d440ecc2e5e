#!/bin/bash
# multi-GPU call: correctness of every exchange / basis combination, the bench line, config 5
# usage: bash scripts/_gpu_multi.sh N
set -u
N="${1:-2}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -ge 8 ]; then export OFB_DIST_QUICK=1; fi
timeout 600 $TR --master-port 29531 tests/dist_gpu_check.py > gpurun_out/r2b_dist_check${N}.log 2>&1; echo "dist check rc=$?"
grep -E "rel_err|lanczos|DIST_CHECK|Error|error" gpurun_out/r2b_dist_check${N}.log | tail -40 | cut -c1-220
timeout 600 $TR --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2b_bench_n${N}.json 2> gpurun_out/r2b_bench_n${N}.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/r2b_bench_n${N}.json; tail -5 gpurun_out/r2b_bench_n${N}.err | cut -c1-300
if [ "$N" -ge 4 ]; then
    timeout 600 $TR --master-port 29533 scripts/run_config5.py > gpurun_out/r2b_config5_n32_${N}gpu.log 2>&1; echo "config5 rc=$?"
    tail -2 gpurun_out/r2b_config5_n32_${N}gpu.log | cut -c1-900
fi
