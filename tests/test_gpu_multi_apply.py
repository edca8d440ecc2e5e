"""K1b (csrc/apply.cu k_apply_multi): H applied to several columns in one launch -- scipy's
_matmat as the block Davidson drives it (linalg/davidson.py:249-256) -- against the oracle's
column-by-column matmat and, bit for bit, against the single-column kernel it shares its
arithmetic with."""
import numpy
import pytest

import helpers
import pauli_oracle as po
from openfermion_b200 import _lib, compiler, hamiltonians
from openfermion_b200.linalg import LinearQubitOperator, QubitDavidson, DavidsonOptions, krylov
from openfermion_b200.plan import PauliPlan

pytestmark = pytest.mark.gpu

CASES = [
    ('hubbard_2x4_n16', lambda: hamiltonians.hubbard_jw(2, 4, 1.0, 4.0, 0.3, 0.2), 16),
    ('molecular_n12', lambda: hamiltonians.molecular_jw(6, seed=1234), 12),
    ('molecular_n16', lambda: hamiltonians.molecular_jw(8, seed=1234), 16),
    ('molecular_bk_n14', lambda: hamiltonians.molecular_jw(7, seed=5, encoding='bk'), 14),
    ('complex_paulis_n13', lambda: hamiltonians.random_pauli_sum(13, 300, 2, True), 13),
    ('tile_sized_n10', lambda: hamiltonians.molecular_jw(5, seed=3), 10),
]


def _columns(n, k, seed):
    rng = numpy.random.default_rng(seed)
    return numpy.asfortranarray(rng.normal(size=(1 << n, k)) + 1j * rng.normal(size=(1 << n, k)))


@pytest.fixture(autouse=True)
def _batched_kernel_on(monkeypatch):
    # K1b is opt-in (measured slower per column than the single-column kernel):
    # these tests keep it correct
    monkeypatch.setenv('OFB_MULTI_APPLY', '1')


@pytest.mark.parametrize('hints', [True, False], ids=['hints', 'plain'])
@pytest.mark.parametrize('name,builder,n', CASES, ids=[c[0] for c in CASES])
def test_matmat_matches_oracle_and_single_column_kernel(name, builder, n, hints):
    op = builder()
    x, z, c = compiler.compile_qubit_operator(op, n)
    plan = PauliPlan(n, x, z, c, hints=hints)
    dim = 1 << n
    for k in (2, 3, 8):
        cols = _columns(n, k, k)
        got = plan.apply_host(cols)
        assert got.shape == (dim, k)
        if n <= 14 or k == 2:
            want = po.matmat(op.terms, n, cols)
            assert helpers.close(got, want, 1e-12)
        single = numpy.stack([plan.apply_host(numpy.ascontiguousarray(cols[:, j])) for j in range(k)], axis=1)
        phased = c * numpy.array([1, -1j, -1, 1j])[hamiltonians.popcount64(x & z) % 4]
        if numpy.any(phased.imag != 0):
            # complex phased coefficients: the single-column kernel runs a segment's real table,
            # then its imaginary table; the batched one alternates per class run (last-bit order)
            assert helpers.close(got, single, 1e-14)
        else:
            assert numpy.array_equal(got, single), 'batched and single-column kernels differ'


def test_device_block_apply_with_leading_dimension():
    n, k = 14, 4
    op = hamiltonians.molecular_jw(7, seed=3)
    lin = LinearQubitOperator(op, n)
    dim = 1 << n
    ld = dim + 512  # columns further apart than their length
    cols = _columns(n, k, 1)
    d_in = _lib.DeviceBuffer(ld * k * 16)
    d_out = _lib.DeviceBuffer(ld * k * 16)
    for j in range(k):
        krylov.upload(numpy.ascontiguousarray(cols[:, j]), d_in.ptr + j * ld * 16)
    lin.plan.apply_device(d_in.ptr, d_out.ptr, k, ld)
    want = po.matmat(op.terms, n, cols)
    for j in range(k):
        got = krylov.download(dim, d_out.ptr + j * ld * 16)
        assert helpers.close(got, want[:, j], 1e-12)


def test_matmat_through_the_linear_operator_api():
    n = 12
    op = hamiltonians.hubbard_jw(2, 3, 1.0, 4.0, 0.2, 0.1)
    lin = LinearQubitOperator(op, n)
    cols = _columns(n, 5, 9)
    assert helpers.close(lin.matmat(cols), po.matmat(op.terms, n, cols), 1e-12)
    assert helpers.close(lin @ cols, po.matmat(op.terms, n, cols), 1e-12)


def test_davidson_with_several_states_uses_the_block_path():
    """n_lowest = 3: every iteration applies H to a block of new directions."""
    n = 12
    op = hamiltonians.molecular_jw(6, seed=1234)
    dense = po.qubit_operator_sparse(op.terms, n).toarray()
    want = numpy.linalg.eigvalsh(dense)[:3]
    numpy.random.seed(3)
    launches = _lib.launch_count()
    success, values, vectors = QubitDavidson(op, n, DavidsonOptions(max_subspace=40, eps=1e-7)).get_lowest_n(
        3, max_iterations=200)
    assert success and _lib.launch_count() > launches
    assert numpy.max(numpy.abs(numpy.sort(values) - want)) < 1e-6
    assert numpy.max(numpy.abs(dense @ vectors - vectors * values)) < 1e-4
