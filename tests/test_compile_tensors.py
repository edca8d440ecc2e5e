"""compiler.compile_interaction_tensors / compile_ladder_products: the array-wise route from an
InteractionOperator's tensors to the projected rows of the Jordan-Wigner matrix must give exactly
the arrays of the reference's route -- get_fermion_operator (conversions.py:117-121, restated in
sparse_tools._tensor_fermion_terms) followed by the per-term compiler -- including what `+=`
drops, products that vanish identically, the term order and the qubit count."""
import numpy
import pytest

from openfermion_b200 import compiler, hamiltonians, ops
from openfermion_b200.linalg import sparse_tools


class _Tensors:
    def __init__(self, n_body_tensors):
        self.n_body_tensors = n_body_tensors


def _both(tensors):
    fop = sparse_tools._tensor_fermion_terms(_Tensors(tensors))
    n = ops.count_qubits(fop)
    want = compiler.compile_fermion_terms(fop.terms, n)
    got_n, got = compiler.compile_interaction_tensors(tensors, None, ops.EQ_TOLERANCE)
    return n, want, got_n, got


def _assert_same(tensors):
    n, want, got_n, got = _both(tensors)
    assert got_n == n
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and numpy.array_equal(a, b)


@pytest.mark.parametrize('n_modes', [1, 2, 3, 5, 8])
def test_random_tensors_equal_the_term_by_term_compiler(n_modes):
    rng = numpy.random.default_rng(n_modes)
    shape2, shape4 = (n_modes,) * 2, (n_modes,) * 4
    h, g = rng.normal(size=shape2), rng.normal(size=shape4)
    _assert_same({(): 0.5, (1, 0): h, (1, 1, 0, 0): g})
    # complex, sparse
    hc = h + 1j * rng.normal(size=shape2)
    gc = g * (rng.random(size=shape4) < 0.3) + 1j * rng.normal(size=shape4) * (rng.random(size=shape4) < 0.1)
    _assert_same({(): 0.0, (1, 0): hc, (1, 1, 0, 0): gc})
    # coefficients below EQ_TOLERANCE are dropped by the reference's +=; an orbital that only
    # occurs in dropped terms does not count as a qubit
    gt = g.copy()
    gt[rng.random(size=shape4) < 0.5] *= 1e-10
    ht = h.copy()
    ht[-1, :] = 0
    ht[:, -1] = 0
    for axis in range(4):
        numpy.moveaxis(gt, axis, 0)[-1] = 1e-11
    _assert_same({(): 1e-12, (1, 0): ht, (1, 1, 0, 0): gt})


def test_edge_cases():
    _assert_same({(): 2.0})
    _assert_same({(1, 0): numpy.zeros((3, 3))})
    _assert_same({(): -1e-9, (1, 0): numpy.eye(2)})
    _assert_same({(1, 1, 0, 0): numpy.ones((2, 2, 2, 2))})  # p == q and r == s vanish identically
    assert compiler.compile_interaction_tensors({(1, 0, 1, 0): numpy.ones((2,) * 4)}, None, 1e-8) is None
    with pytest.raises(ValueError):
        compiler.compile_interaction_tensors({(1, 0): numpy.eye(3)}, 2, 1e-8)


def test_lih_integrals():
    import os
    data = numpy.load(os.path.join(os.path.dirname(__file__), 'golden', 'lih_sto3g.npz'))
    ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                             data['two_body_integrals'])
    n, want, got_n, got = _both(ham.n_body_tensors)
    assert n == got_n == 12 and len(want[0]) == 1501
    for a, b in zip(got, want):
        assert numpy.array_equal(a, b)


def test_ladder_products_against_the_scalar_compiler_on_arbitrary_patterns():
    rng = numpy.random.default_rng(9)
    n = 7
    for kinds in ((1,), (0,), (1, 0), (0, 1), (1, 1, 0), (1, 0, 1, 0), (0, 0, 1, 1), (1, 1, 1, 0, 0, 0)):
        modes = rng.integers(0, n, size=(200, len(kinds)))
        coeff = rng.normal(size=200) + 1j * rng.normal(size=200)
        coeff[rng.random(200) < 0.1] = 0.0
        got = compiler.compile_ladder_products(modes, kinds, coeff, n)
        want = [[], [], [], [], []]
        for row, c in zip(modes, coeff):  # one term at a time: duplicates must not merge
            one = compiler.compile_fermion_terms({tuple(zip(row.tolist(), kinds)): c}, n)
            for k in range(5):
                want[k].extend(one[k].tolist())
        for a, b in zip(got, want):
            assert a.tolist() == b
