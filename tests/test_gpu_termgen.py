"""N3 on the device (csrc/termgen.cu): the GPU expansion + merge of ladder-operator products
against the host builder (openfermion_b200/hamiltonians.py, itself validated against the
reference's jordan_wigner / bravyi_kitaev in tests/test_transforms.py and the golden operators).
Same strings in the same (x, z) order; coefficients equal bit for bit because both sum the
expanded strings in insertion order."""
import time

import numpy
import pytest

import helpers
from openfermion_b200 import hamiltonians, transforms
from openfermion_b200.ops import FermionOperator

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert a.n_qubits == b.n_qubits
    assert numpy.array_equal(a.xq, b.xq) and numpy.array_equal(a.zq, b.zq)
    assert numpy.array_equal(a.coeff, b.coeff), numpy.max(numpy.abs(a.coeff - b.coeff))


@pytest.mark.parametrize('encoding', ['jw', 'bk'])
@pytest.mark.parametrize('n_orbitals', [2, 3, 5, 8])
def test_molecular_tables(n_orbitals, encoding):
    host = hamiltonians.molecular_jw(n_orbitals, seed=7, encoding=encoding)
    dev = hamiltonians.molecular_jw(n_orbitals, seed=7, encoding=encoding, device=True)
    _same(dev, host)
    dyadic = hamiltonians.molecular_jw(n_orbitals, seed=7, dyadic_bits=16, encoding=encoding, device=True)
    _same(dyadic, hamiltonians.molecular_jw(n_orbitals, seed=7, dyadic_bits=16, encoding=encoding))


@pytest.mark.parametrize('encoding', ['jw', 'bk'])
def test_hubbard_tables(encoding):
    for dims in ((2, 2), (2, 4), (3, 4), (4, 4)):
        args = dims + (1.0, 4.0, 0.3, 0.2)
        _same(hamiltonians.hubbard_jw(*args, encoding=encoding, device=True),
              hamiltonians.hubbard_jw(*args, encoding=encoding))


def test_more_than_32_qubits_uses_the_two_pass_sort():
    _same(hamiltonians.hubbard_jw(4, 5, device=True), hamiltonians.hubbard_jw(4, 5))  # 40 qubits


def test_fermion_operator_with_complex_coefficients_and_cancellations():
    op = FermionOperator()
    rng = numpy.random.default_rng(5)
    for _ in range(60):
        k = int(rng.integers(1, 5))
        term = tuple((int(rng.integers(0, 7)), int(rng.integers(0, 2))) for _ in range(k))
        op += FermionOperator(term, complex(rng.normal(), rng.normal()))
    op += FermionOperator(((1, 1), (2, 0)), 0.75) + FermionOperator(((1, 1), (2, 0)), -0.75)  # cancels
    op += FermionOperator((), 0.5)
    _same(transforms.jordan_wigner(op, device=True), transforms.jordan_wigner(op))
    _same(transforms.bravyi_kitaev(op, 9, device=True), transforms.bravyi_kitaev(op, 9))


def test_golden_reference_operator_and_timing_at_28_qubits():
    data = helpers.golden('qubit_cases.npz')
    ref = helpers.qubit_case_operator(helpers.Case(data, 'molecular_8_jw'))
    ours = hamiltonians.molecular_jw(4, seed=1234, device=True).terms
    assert set(ours) == set(ref.terms)
    assert max(abs(ours[k] - ref.terms[k]) for k in ours) < 1e-13
    tensors = hamiltonians.random_interaction_tensors(14, 1234)
    t0 = time.perf_counter()
    dev = hamiltonians.interaction_operator_jw(*tensors, device=True)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host = hamiltonians.interaction_operator_jw(*tensors)
    t_host = time.perf_counter() - t0
    _same(dev, host)
    assert len(dev) == 88495
    print('28-qubit term table: device %.3f s, host %.3f s' % (t_dev, t_host))
