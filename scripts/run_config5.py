#!/usr/bin/env python
"""BASELINE.json config 5 end to end: Fermi-Hubbard 4x4 (32 qubits), state sharded by the
high qubits over the GPUs of one node, device-resident Lanczos ground state.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_config5.py
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.distributed import ShardedPauliOperator, init_nccl, lanczos_ground_state  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--x', type=int, default=4)
    ap.add_argument('--y', type=int, default=4)
    ap.add_argument('--mode', default='ce')
    ap.add_argument('--basis', default='symmetry', choices=['symmetry', 'natural'])
    ap.add_argument('--order', default='locality', choices=['locality', 'natural'])
    ap.add_argument('--tol', type=float, default=1e-8)
    ap.add_argument('--krylov', type=int, default=120)
    args = ap.parse_args()
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    dist = init_nccl(local)
    _lib.call('ofb_set_device', ctypes.c_int(local))
    op = hamiltonians.hubbard_jw(args.x, args.y)
    sop = ShardedPauliOperator(op, op.n_qubits, dist=dist, mode=args.mode, basis=args.basis,
                               local_order=args.order if args.basis == 'symmetry' else 'natural')
    stats = {}
    import torch
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    energy, vec, residual = lanczos_ground_state(sop, dist, tol=args.tol, krylov_dim=args.krylov,
                                                 stats=stats)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({'config': 'Fermi-Hubbard %dx%d (%d qubits) sharded Lanczos' % (args.x, args.y, op.n_qubits),
                          'n_gpus': world, 'exchange': args.mode,
                          'shard_bits': sop.basis.describe(sop.layout.shard_bits),
                          'peer_shards_per_matvec': len(sop.layout.remote_patterns), 'E0': energy, 'residual': residual,
                          'matvecs': stats.get('matvecs'), 'restarts': stats.get('restarts'),
                          'seconds': dt, 'seconds_per_matvec_incl_vector_ops': dt / max(stats.get('matvecs', 1), 1)}))
    sop.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
