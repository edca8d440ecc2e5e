"""The support-hint kernel variant (k_apply_tile<..., COND = true>, ofb_plan_set_support,
openfermion_b200/support.py; the default since round 2) against the plain kernel on the full
term table: same H|psi>, expectation and sharded partial applies with and without hints; the
sector kernel on the orbit-reduced table; stored-basis Lanczos against the two-pass scheme."""
import os

import numpy
import pytest

import helpers
from openfermion_b200 import compiler, hamiltonians

pytestmark = pytest.mark.gpu

CASES = [
    ('hubbard_2x2_n8_small_state_kernel', lambda: hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2), 8),
    ('molecular_n6_small_state_kernel', lambda: hamiltonians.molecular_jw(3, seed=2), 6),
    ('hubbard_2x4_n16', lambda: hamiltonians.hubbard_jw(2, 4), 16),
    ('hubbard_2x5_mu_h_n20', lambda: hamiltonians.hubbard_jw(2, 5, 1.0, 4.0, 0.3, 0.2), 20),
    ('hubbard_2x4_bk_n16', lambda: hamiltonians.hubbard_jw(2, 4, encoding='bk'), 16),
    ('molecular_n12', lambda: hamiltonians.molecular_jw(6, seed=1234), 12),
    ('molecular_n16', lambda: hamiltonians.molecular_jw(8, seed=1234), 16),
    ('molecular_bk_n14', lambda: hamiltonians.molecular_jw(7, seed=5, encoding='bk'), 14),
    ('complex_paulis_n13', lambda: hamiltonians.random_pauli_sum(13, 200, 2, True), 13),
]


def _state(n, seed):
    rng = numpy.random.default_rng(seed)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / numpy.linalg.norm(v)


@pytest.mark.parametrize('reduce', [False, True], ids=['hints', 'hints+reduced'])
@pytest.mark.parametrize('name,builder,n', CASES, ids=[c[0] for c in CASES])
def test_hints_do_not_change_the_result(name, builder, n, reduce):
    from openfermion_b200.plan import PauliPlan
    op = builder()
    x, z, c = compiler.compile_qubit_operator(op, n)
    plain = PauliPlan(n, x, z, c, hints=False)
    hinted = PauliPlan(n, x, z, c, hints=False)
    hinted.set_support_hints(x, z, c, reduce=reduce)
    v = _state(n, 3)
    want = plain.apply_host(v)
    got = hinted.apply_host(v)
    assert helpers.close(got, want, 1e-13)
    assert abs(hinted.expectation_host(v) - plain.expectation_host(v)) <= 1e-12 * numpy.sum(numpy.abs(c))
    if 'hubbard' in name or 'molecular' in name:
        assert hinted.support_fraction < 0.6
        if reduce:
            assert hinted.n_reduced_terms < plain.n_terms
    # diagonal and CSC keep the full table
    if n <= 14 and not numpy.any(c.imag):
        assert numpy.array_equal(hinted.diagonal_host(), plain.diagonal_host())


def test_hints_with_shards_and_accumulate():
    """index_base bits take part in the parity of a condition: virtual ranks on one GPU."""
    from openfermion_b200 import _lib
    from openfermion_b200.distributed import ShardLayout
    from openfermion_b200.linalg import krylov
    from openfermion_b200.plan import PauliPlan
    from openfermion_b200.shard_basis import ShardBasis
    n, world = 16, 8
    op = hamiltonians.hubbard_jw(2, 4)
    x, z, c = compiler.compile_qubit_operator(op, n)
    for kind in ('natural', 'symmetry'):
        basis = ShardBasis.natural(n) if kind == 'natural' else \
            ShardBasis.for_sharding(n, 3, x, local_order='locality')
        xb, zb, cb = basis.transform_terms(x, z, c)
        plain = PauliPlan(n, xb, zb, cb, hints=False)
        hinted = PauliPlan(n, xb, zb, cb, hints=False)
        hinted.set_support_hints(xb, zb, cb)
        layout = ShardLayout(n, world, xb)
        v = basis.from_natural(_state(n, 5))
        d_in = _lib.to_device(v)
        outs = []
        for plan in (plain, hinted):
            d_out = _lib.DeviceBuffer(v.nbytes)
            ld = layout.local_dim
            for rank in range(world):
                first = True
                for x_hi, g0, g1 in layout.patterns:
                    src = d_in.ptr + layout.peer(rank, x_hi) * ld * 16
                    plan.apply_groups_device(src, d_out.ptr + rank * ld * 16, layout.local_bits,
                                             layout.index_base(rank), g0, g1, 0 if first else 1)
                    first = False
            outs.append(krylov.download(1 << n, d_out.ptr))
        assert helpers.close(outs[1], outs[0], 1e-13)


@pytest.mark.parametrize('kind', ['sz_number', 'number'])
def test_sector_kernel_on_the_reduced_table(kind, monkeypatch):
    """SectorLinearQubitOperator.apply_plan with support hints (the default): the unchanged sector kernel
    on the orbit-reduced table (only the conditions sector membership implies)."""
    from openfermion_b200.linalg import SectorLinearQubitOperator, expectation
    for op, n in ((hamiltonians.hubbard_jw(2, 4, 1.0, 4.0, 0.3, 0.2), 16), (hamiltonians.molecular_jw(7, seed=4), 14)):
        kwargs = dict(n_electrons=6, sz_value=0.0) if kind == 'sz_number' else dict(n_electrons=5)
        monkeypatch.setenv('OFB_SUPPORT_HINTS', '0')
        plain = SectorLinearQubitOperator(op, n, **kwargs)
        rng = numpy.random.default_rng(1)
        v = rng.normal(size=plain.shape[0]) + 1j * rng.normal(size=plain.shape[0])
        want = plain * v
        e_want = expectation(plain, v)
        monkeypatch.setenv('OFB_SUPPORT_HINTS', '1')
        reduced = SectorLinearQubitOperator(op, n, **kwargs)
        got = reduced * v
        assert helpers.close(got, want, 1e-13)
        assert abs(expectation(reduced, v) - e_want) <= 1e-12 * abs(e_want)
        if kind == 'sz_number':
            assert reduced.apply_plan.n_terms < plain.plan.n_terms


def test_stored_basis_lanczos_gives_the_same_ground_state(monkeypatch):
    """Stored-basis Lanczos (linalg/krylov.py:_lanczos_pass_stored, the default): the Krylov vectors stay in HBM and
    the Ritz vector is one pass over them -- about half the matrix-vector products, same result."""
    from openfermion_b200.linalg import LinearQubitOperator, SectorLinearQubitOperator, get_ground_state, krylov
    for build in (lambda: LinearQubitOperator(hamiltonians.hubbard_jw(2, 4), 16),
                  lambda: LinearQubitOperator(hamiltonians.molecular_jw(6, seed=1234), 12),
                  lambda: SectorLinearQubitOperator(hamiltonians.hubbard_jw(3, 4), 24, n_electrons=12, sz_value=0.0)):
        monkeypatch.setenv('OFB_LANCZOS_STORE', '0')
        op = build()
        stats_two_pass, stats_stored = {}, {}
        dev = krylov.as_device_operator(op)
        e_two, vec, res_two = krylov.lanczos_lowest(dev, stats=stats_two_pass)
        monkeypatch.setenv('OFB_LANCZOS_STORE', '1')
        e_one, vec1, res_one = krylov.lanczos_lowest(dev, stats=stats_stored)
        assert abs(e_one - e_two) <= 1e-10 * max(1.0, abs(e_two))
        assert res_one <= 1e-8 * max(1.0, abs(e_one))
        assert stats_stored['matvecs'] < 0.7 * stats_two_pass['matvecs']
        energy, state = get_ground_state(op)
        assert abs(energy - e_two) <= 1e-10 * max(1.0, abs(e_two))
