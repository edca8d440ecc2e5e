"""Multi-GPU check, run under torchrun on N GPUs (not collected by pytest):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_check.py
Compares the sharded H|psi> (both data paths) and the distributed Lanczos energy with the
single-GPU path on rank 0 at sizes that fit one GPU."""
import ctypes
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.distributed import ShardedPauliOperator, TorchComm, lanczos_ground_state  # noqa: E402
from openfermion_b200.linalg import LinearQubitOperator, get_ground_state, krylov  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    from openfermion_b200.distributed import init_nccl
    init_nccl(local)
    _lib.call('ofb_set_device', ctypes.c_int(local))
    ok = True
    quick = bool(os.environ.get('OFB_DIST_QUICK'))  # fewer mode/basis combinations (GPU-minutes)
    for name, op, n in (('hubbard_2x4', hamiltonians.hubbard_jw(2, 4), 16),
                        ('molecular_12', hamiltonians.molecular_jw(6, seed=1234), 12),
                        ('dense_complex_14', hamiltonians.random_pauli_sum(14, 300, 3, True), 14)):
        rng = numpy.random.default_rng(1)
        full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        combos = (('nccl', 'natural', 'natural'), ('p2p', 'natural', 'natural'), ('ce', 'natural', 'natural'),
                  ('ce', 'symmetry', 'natural'), ('nccl', 'symmetry', 'natural'), ('p2p', 'symmetry', 'natural'),
                  ('ce', 'symmetry', 'locality'), ('p2p', 'symmetry', 'locality'))
        if quick:
            combos = (('ce', 'natural', 'natural'), ('ce', 'symmetry', 'natural'), ('ce', 'symmetry', 'locality'))
        for mode, basis, order in combos:
            sop = ShardedPauliOperator(op, n, dist=dist, mode=mode, basis=basis, local_order=order,
                                       max_staging=1 if mode in ('nccl', 'ce') else None)
            ld = sop.dim
            vin, vout = sop.new_vector(), sop.new_vector()
            krylov.upload(numpy.ascontiguousarray(sop.shard_of(full)), vin.ptr)
            if mode in ('p2p', 'ce'):
                sop.register_peer_vectors([vin])
                TorchComm(dist, 'cuda:%d' % local).barrier()
            sop.matvec(vin, vout)
            torch.cuda.synchronize()
            got = krylov.download(ld, vout.ptr)
            gathered = [None] * world
            dist.all_gather_object(gathered, got)
            if rank == 0:
                want = LinearQubitOperator(op, n) * full
                result = sop.basis.to_natural(numpy.concatenate(gathered))
                err = numpy.max(numpy.abs(result - want)) / numpy.max(numpy.abs(want))
                print('%s %s basis=%s order=%s world=%d rel_err=%.2e patterns=%d' % (
                    name, mode, basis, order, world, err, len(sop.layout.remote_patterns)))
                ok = ok and err < 1e-12
            dist.barrier()
            sop.close()
    hub = hamiltonians.hubbard_jw(2, 4)
    e1 = None
    if rank == 0:
        e1, _ = get_ground_state(LinearQubitOperator(hub, 16))
    for mode, basis, order in ((('ce', 'natural', 'natural'), ('ce', 'symmetry', 'locality')) if quick else
                               (('nccl', 'natural', 'natural'), ('ce', 'natural', 'natural'),
                                ('p2p', 'natural', 'natural'), ('ce', 'symmetry', 'natural'),
                                ('ce', 'symmetry', 'locality'))):
        sop = ShardedPauliOperator(hub, 16, dist=dist, mode=mode, basis=basis, local_order=order)
        # two consecutive solves on ONE operator (the first solve's buffers are
        # unregistered from the peers before they are freed, so the second cannot hit stale maps)
        for attempt in range(2):
            stats = {}
            energy, vec, residual = lanczos_ground_state(sop, dist, stats=stats, seed=1234 + attempt)
            vec.free()
            if rank == 0:
                print('lanczos %s/%s/%s world=%d solve %d E=%.12f single-GPU E=%.12f diff=%.2e residual=%.2e matvecs=%d'
                      % (mode, basis, order, world, attempt, energy, e1, abs(energy - e1), residual,
                         stats.get('matvecs', 0)))
                ok = ok and abs(energy - e1) < 1e-10
        sop.close()
    if rank == 0:
        print('DIST_CHECK', 'PASS' if ok else 'FAIL')
    if os.environ.get('OFB_DIST_BENCH_QUBITS'):
        compare_bases(int(os.environ['OFB_DIST_BENCH_QUBITS']), rank, world, local)
    dist.destroy_process_group()


def compare_bases(n, rank, world, local, steps=5, warmup=3):
    """Same Hubbard H|psi> sharded by the top qubits and by conserved parities, timed back to
    back in one process group (CUDA events, max over ranks); one JSON line per basis."""
    import json
    dims = {16: (2, 4), 20: (2, 5), 24: (3, 4), 28: (2, 7), 30: (3, 5), 32: (4, 4)}[n]
    model = hamiltonians.hubbard_jw(*dims)
    comm = TorchComm(dist, 'cuda:%d' % local)
    for basis, local_order in (('natural', 'natural'), ('symmetry', 'natural'), ('symmetry', 'locality')):
        sop = ShardedPauliOperator(model, n, dist=dist, mode='ce', basis=basis, local_order=local_order)
        vin, vout = sop.new_vector(), sop.new_vector()
        krylov.fill_random(sop.dim, 42 + rank, False, vin.ptr)
        exchanging = bool(sop.layout.remote_patterns)
        if exchanging:
            sop.register_peer_vectors([vin])

        def step():
            if exchanging:
                comm.barrier()
            sop.matvec(vin, vout)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e = krylov.gdot(sop.dim, vin.ptr, vout.ptr, comm)
        if rank == 0:
            print(json.dumps({'workload': 'Fermi-Hubbard %dx%d JW, %d qubits' % (dims + (n,)), 'n_gpus': world,
                              'basis': basis, 'local_qubit_order': local_order, 'shard_bits': sop.basis.describe(sop.layout.shard_bits),
                              'peer_shards_per_matvec': len(sop.layout.remote_patterns),
                              'ms_per_matvec': float(t.item()),
                              'term_amp_per_s': sop.n_terms * float(1 << n) / (float(t.item()) / 1e3),
                              'vHv': [e.real, e.imag]}))
        sop.close()
        torch.cuda.synchronize()
        # free the 2^n-sized buffers now: the operator holds reference cycles (bound methods)
        for buf in [vin, vout] + list(sop._staging):
            buf.free()
        del vin, vout, sop
        import gc
        gc.collect()
        dist.barrier()


if __name__ == '__main__':
    main()
