"""CPU, gloo, world_size 2 and 4: the host logic of the sharded H|psi> (pattern table,
peers, staging schedule, accumulate flags, reductions) with an injected numpy local-apply
that restates the gather form; the result is compared with the oracle's unsharded matvec."""
import os
import socket
import sys

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pauli_oracle as po
from openfermion_b200 import hamiltonians
from openfermion_b200.distributed import ShardLayout, ShardedPauliOperator, TorchComm
from openfermion_b200.ops import QubitOperator


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _HostVec:
    def __init__(self, n):
        self.array = numpy.zeros(n, dtype=numpy.complex128)
        self.ptr = id(self)


def _make_operator(kind):
    if kind == 'hubbard':
        return hamiltonians.hubbard_jw(2, 2), 8
    if kind == 'complex':
        return hamiltonians.random_pauli_sum(7, 40, 5, True), 7
    op = QubitOperator('X0 Y1', 0.5) + QubitOperator('Z0 Z3', -1.5) + QubitOperator('Y2 X5', 2j) \
        + QubitOperator((), 0.25)
    return op, 6


def _worker(rank, world, port, kind, depth, out_dir, basis='natural', local_order='natural'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        operator, n = _make_operator(kind)
        applied = []
        ctx = {}

        def local_apply(vin, vout, g0, g1, accumulate):
            # numpy restatement of the gather form on this rank's shard, over the operator's
            # own (possibly relabelled) tables
            x, z, c = ctx['tables']
            layout, j, jg, group_of = ctx['layout'], ctx['j'], ctx['jg'], ctx['group_of']
            ld = layout.local_dim
            applied.append((g0, g1, accumulate))
            out = vout.array
            if not accumulate:
                out[:] = 0
            for t in range(len(c)):
                if not (g0 <= group_of[t] < g1):
                    continue
                y = bin(int(x[t] & z[t])).count('1')
                phase = c[t] * (-1j) ** y
                sign = 1.0 - 2.0 * po.bit_parity(jg & z[t]).astype(float)
                out += phase * sign * vin.array[(j ^ (x[t] & numpy.uint64(ld - 1))).astype(numpy.int64)]

        op = ShardedPauliOperator(
            operator, n, dist=dist, mode='nccl', max_staging=depth, local_apply=local_apply,
            vector_factory=lambda: _HostVec(1 << (n - (world.bit_length() - 1))),
            tensor_of=lambda v: torch.from_numpy(v.array.view(numpy.float64)), basis=basis,
            local_order=local_order)
        layout = op.layout
        ld = layout.local_dim
        j = numpy.arange(ld, dtype=numpy.uint64)
        ctx.update(tables=op._tables, layout=layout, j=j, jg=j | numpy.uint64(rank << layout.local_bits),
                   group_of=numpy.searchsorted(layout.group_x, op._tables[0]))
        rng = numpy.random.default_rng(3)
        full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        vin, vout = op.new_vector(), op.new_vector()
        vin.array[:] = op.shard_of(full)
        op.matvec(vin, vout)
        want = op.shard_of(po.matvec(operator.terms, n, full))
        err = float(numpy.max(numpy.abs(vout.array - want)))
        # every group applied exactly once, first call overwrites, later calls accumulate
        covered = sorted((a, b) for a, b, _ in applied if b > a)
        ok_cover = covered == sorted((g0, g1) for _, g0, g1 in layout.patterns)
        ok_flags = [acc for _, _, acc in applied] == [False] + [True] * (len(applied) - 1)
        comm = TorchComm(dist, 'cpu')
        total = comm.allsum(numpy.vdot(vin.array, vout.array))
        want_total = numpy.vdot(full, po.matvec(operator.terms, n, full))
        # the shards of all ranks tile the natural index space exactly once
        seen = numpy.zeros(1 << n, dtype=numpy.int64)
        for r in range(world):
            seen[op.natural_indices(r).astype(numpy.int64)] += 1
        numpy.save(os.path.join(out_dir, 'r%d.npy' % rank),
                   numpy.array([err, float(ok_cover), float(ok_flags), abs(total - want_total),
                                len(layout.remote_patterns), float(numpy.all(seen == 1))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,kind,depth', [(2, 'hubbard', None), (2, 'complex', 1), (4, 'hubbard', 1),
                                              (4, 'complex', None), (2, 'small', None), (4, 'small', 2)])
def test_sharded_matvec_over_gloo(world, kind, depth, tmp_path):
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, kind, depth, str(tmp_path)), nprocs=world,
                       join=True, start_method='fork')
    for rank in range(world):
        err, ok_cover, ok_flags, dot_err, n_remote, tiled = numpy.load(tmp_path / ('r%d.npy' % rank))
        assert err < 1e-12, (rank, err)
        assert ok_cover == 1.0 and ok_flags == 1.0 and tiled == 1.0
        assert dot_err < 1e-11


@pytest.mark.parametrize('local_order', ['natural', 'locality'])
@pytest.mark.parametrize('world,kind,remote', [(2, 'hubbard', 0), (4, 'hubbard', 0), (8, 'hubbard', 1),
                                               (4, 'complex', None), (2, 'small', 0)])
def test_symmetry_basis_sharded_matvec_over_gloo(world, kind, remote, local_order, tmp_path):
    """Sharding by conserved parities (shard_basis.py): same H|psi> after relabelling, and the
    spin-parity bits of a Hubbard model need no peer shard at all."""
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, kind, None, str(tmp_path), 'symmetry', local_order),
                       nprocs=world,
                       join=True, start_method='fork')
    for rank in range(world):
        err, ok_cover, ok_flags, dot_err, n_remote, tiled = numpy.load(tmp_path / ('r%d.npy' % rank))
        assert err < 1e-12, (rank, err)
        assert ok_cover == 1.0 and ok_flags == 1.0 and tiled == 1.0
        assert dot_err < 1e-11
        if remote is not None:
            assert n_remote == remote


def test_shard_layout_patterns():
    op = hamiltonians.hubbard_jw(4, 4)
    x, _ = op.index_masks(32)
    # 4/65 groups cross at k=1 (1 pattern), 8/65 at k=2 (2), 11/65 at k=3 (4)
    for world, crossing, patterns in ((2, 4, 1), (4, 8, 2), (8, 11, 4)):
        layout = ShardLayout(32, world, x)
        assert len(layout.group_x) == 65
        assert len(layout.remote_patterns) == patterns
        assert sum(g1 - g0 for _, g0, g1 in layout.remote_patterns) == crossing
        assert layout.patterns[0][0] == 0
        covered = sorted((g0, g1) for _, g0, g1 in layout.patterns)
        assert covered[0][0] == 0 and covered[-1][1] == 65
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
        assert layout.exchange_bytes_per_matvec() == patterns * (2**32 // world) * 16
    with pytest.raises(ValueError):
        ShardLayout(4, 3, x[:1])
    one = ShardLayout(5, 1, numpy.array([0, 3, 17], dtype=numpy.uint64))
    assert one.remote_patterns == [] and one.local_pattern == (0, 0, 3)
    assert ShardLayout(6, 4, numpy.array([1], dtype=numpy.uint64)).peer(2, 3) == 1


def _observable_worker(rank, world, port, out_dir):
    """An observable applied on shards cut by ANOTHER operator's symmetry basis."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from openfermion_b200.compiler import compile_qubit_operator
        from openfermion_b200.shard_basis import ShardBasis
        n = 8
        hamiltonian = hamiltonians.hubbard_jw(2, 2)
        basis = ShardBasis.for_sharding(n, world.bit_length() - 1, compile_qubit_operator(hamiltonian, n)[0],
                                        local_order='locality')
        observable = hamiltonians.random_pauli_sum(n, 25, 9, True)  # shares no symmetry with H
        ctx = {}

        def local_apply(vin, vout, g0, g1, accumulate):
            x, z, c = ctx['tables']
            layout, j, jg, group_of = ctx['layout'], ctx['j'], ctx['jg'], ctx['group_of']
            out = vout.array
            if not accumulate:
                out[:] = 0
            for t in range(len(c)):
                if g0 <= group_of[t] < g1:
                    phase = c[t] * (-1j) ** bin(int(x[t] & z[t])).count('1')
                    sign = 1.0 - 2.0 * po.bit_parity(jg & z[t]).astype(float)
                    out += phase * sign * vin.array[(j ^ (x[t] & numpy.uint64(layout.local_dim - 1))).astype(numpy.int64)]

        op = ShardedPauliOperator(
            observable, n, dist=dist, mode='nccl', local_apply=local_apply,
            vector_factory=lambda: _HostVec(1 << (n - (world.bit_length() - 1))),
            tensor_of=lambda v: torch.from_numpy(v.array.view(numpy.float64)), basis=basis)
        layout = op.layout
        j = numpy.arange(layout.local_dim, dtype=numpy.uint64)
        ctx.update(tables=op._tables, layout=layout, j=j, jg=j | numpy.uint64(rank << layout.local_bits),
                   group_of=numpy.searchsorted(layout.group_x, op._tables[0]))
        rng = numpy.random.default_rng(6)
        full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        vin, vout = op.new_vector(), op.new_vector()
        vin.array[:] = op.shard_of(full)
        op.matvec(vin, vout)
        want_full = po.matvec(observable.terms, n, full)
        err = float(numpy.max(numpy.abs(vout.array - op.shard_of(want_full))))
        total = TorchComm(dist, 'cpu').allsum(numpy.vdot(vin.array, vout.array))
        numpy.save(os.path.join(out_dir, 'o%d.npy' % rank),
                   numpy.array([err, abs(total - numpy.vdot(full, want_full))]))
    finally:
        dist.destroy_process_group()


def test_observable_in_the_basis_of_another_operator(tmp_path):
    world = 4
    mp.start_processes(_observable_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world,
                       join=True, start_method='fork')
    for rank in range(world):
        err, dot_err = numpy.load(tmp_path / ('o%d.npy' % rank))
        assert err < 1e-12 and dot_err < 1e-11
