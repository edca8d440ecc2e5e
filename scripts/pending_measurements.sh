#!/bin/bash
# Measurements that were prepared on CPU after the round-1 GPU budget was spent; run each block
# under gpurun with the stated GPU count (see DESIGN.md section 4 K6 "Qubit order inside a shard").
#
#   gpurun --timeout 300 -- 'bash scripts/pending_measurements.sh one'
#   gpurun --timeout 300 -- 'bash scripts/pending_measurements.sh hints'
#   gpurun --timeout 600 -- 'bash scripts/pending_measurements.sh ncu'
#   gpurun --gpus 2 --timeout 200 -- 'bash scripts/pending_measurements.sh multi 2'
#   gpurun --gpus 8 --timeout 300 -- 'bash scripts/pending_measurements.sh multi 8'
set -u
mkdir -p gpurun_out
case "${1:-one}" in
one)
    # the whole GPU suite (includes the reference-ladder file added after the last GPU run)
    python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pending.log 2>&1
    tail -3 gpurun_out/pytest_gpu_pending.log
    ;;
hints)
    # experimental support-hint kernel (openfermion_b200/support.py): correctness, then timing
    OFB_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_support_hints.py -m gpu -x -q 2>&1 | tail -3
    for w in hub30 mol26; do
        python scripts/quick_apply_bench.py $w
        OFB_SUPPORT_HINTS=hints python scripts/quick_apply_bench.py $w
        OFB_SUPPORT_HINTS=1 python scripts/quick_apply_bench.py $w
    done
    python scripts/sector_bench.py hub44_half mol28
    OFB_SUPPORT_HINTS=1 python scripts/sector_bench.py hub44_half mol28
    # stored-basis Lanczos (OFB_LANCZOS_STORE=1): time to the ground state, sector cases
    python scripts/sector_bench.py hub44_half --lanczos
    OFB_LANCZOS_STORE=1 python scripts/sector_bench.py hub44_half --lanczos
    ;;
ncu)
    # one full capture of the hinted kernel per workload (only after the plain runs above exited 0)
    for w in hub30 mol26; do
        OFB_SUPPORT_HINTS=1 ncu --set full --clock-control none --import-source on -k regex:k_apply_tile \
            -s 2 -c 1 -o "gpurun_out/k_apply_hints_${w}" python scripts/quick_apply_bench.py $w \
            > "gpurun_out/ncu_hints_${w}.log" 2>&1
        ncu -i "gpurun_out/k_apply_hints_${w}.ncu-rep" --page raw --csv > "gpurun_out/k_apply_hints_${w}_raw.csv" 2>/dev/null
    done
    ;;
multi)
    N="${2:-2}"
    # correctness of every basis / exchange combination, then the 32-qubit Hubbard H|psi> with
    # top-qubit shards, parity shards, and parity shards + locality-ordered local qubits
    OFB_DIST_QUICK=1 OFB_DIST_BENCH_QUBITS=32 timeout 200 python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 tests/dist_gpu_check.py \
        > "gpurun_out/dist_check${N}_locality.log" 2>&1
    tail -8 "gpurun_out/dist_check${N}_locality.log"
    # config 5 (32-qubit Lanczos ground state) in the new basis
    if [ "$N" = 8 ]; then
        timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
            --master-port 29532 scripts/run_config5.py > gpurun_out/config5_n32_8gpu_symmetry.log 2>&1
        tail -2 gpurun_out/config5_n32_8gpu_symmetry.log
    fi
    ;;
esac
