"""Array-form Jordan-Wigner / Bravyi-Kitaev transforms (openfermion_b200/transforms.py) against
the reference's own jordan_wigner / bravyi_kitaev outputs stored in tests/golden/sector_cases.npz
(oracle/make_golden.py)."""
import numpy
import pytest

from helpers import fermion_operator_from, golden, qubit_operator_from
from openfermion_b200 import hamiltonians, ops, transforms

DATA = golden('sector_cases.npz')
OPERATORS = [str(n) for n in DATA['names']]


def _assert_same_operator(mask_operator, reference_terms, tol=1e-13):
    got = mask_operator.terms
    assert set(got) == set(reference_terms)
    assert max(abs(got[k] - reference_terms[k]) for k in reference_terms) <= tol


@pytest.mark.parametrize('name', OPERATORS)
def test_jordan_wigner_and_bravyi_kitaev_match_reference(name):
    fop = fermion_operator_from(DATA[name + '/fermion_terms'], DATA[name + '/fermion_coeffs'],
                                DATA[name + '/fermion_coeff_is_complex'])
    jw_ref = qubit_operator_from(DATA[name + '/qubit_terms'], DATA[name + '/qubit_coeffs'])
    bk_ref = qubit_operator_from(DATA[name + '/bk_terms'], DATA[name + '/bk_coeffs'])
    _assert_same_operator(transforms.jordan_wigner(fop), jw_ref.terms)
    _assert_same_operator(transforms.bravyi_kitaev(fop), bk_ref.terms)


def test_tensor_and_diagonal_coulomb_inputs():
    constant, one, two = hamiltonians.random_interaction_tensors(3, seed=5)
    tensor = hamiltonians.InteractionTensor(constant, one, two)
    # molecular_6 of the golden file is random_interaction_operator(3, expand_spin=True, seed=5)
    jw_ref = qubit_operator_from(DATA['molecular_6/qubit_terms'], DATA['molecular_6/qubit_coeffs'])
    _assert_same_operator(transforms.jordan_wigner(tensor), jw_ref.terms)
    bk_ref = qubit_operator_from(DATA['molecular_6/bk_terms'], DATA['molecular_6/bk_coeffs'])
    _assert_same_operator(transforms.bravyi_kitaev(tensor), bk_ref.terms)
    # n_p n_q = (1 - Z_p)(1 - Z_q) / 4: a diagonal Coulomb operator maps to Z strings only
    dch = hamiltonians.DiagonalCoulomb(numpy.zeros((3, 3)), numpy.array([[0, 1.0, 0], [1.0, 0, 0], [0, 0, 0]]), 0.5)
    image = transforms.jordan_wigner(dch)
    assert all(action == 'Z' for term in image.terms for _, action in term)
    assert abs(image.terms[()] - 1.0) < 1e-14 and abs(image.terms[((0, 'Z'), (1, 'Z'))] - 0.5) < 1e-14


def test_known_small_transforms_and_errors():
    # jordan_wigner_test.py: a_3^ -> 1/2 Z0 Z1 Z2 (X3 - i Y3)
    image = transforms.jordan_wigner(ops.FermionOperator('3^')).terms
    assert image == {((0, 'Z'), (1, 'Z'), (2, 'Z'), (3, 'X')): 0.5,
                     ((0, 'Z'), (1, 'Z'), (2, 'Z'), (3, 'Y')): -0.5j}
    # number operator: (1 - Z_2) / 2 in both encodings' JW form
    number = transforms.jordan_wigner(ops.FermionOperator('2^ 2')).terms
    assert number == {(): 0.5, ((2, 'Z'),): -0.5}
    # bravyi_kitaev_test.py: a_0^ a_0 on one mode is (1 - Z_0) / 2 as well
    assert transforms.bravyi_kitaev(ops.FermionOperator('0^ 0')).terms == {(): 0.5, ((0, 'Z'),): -0.5}
    with pytest.raises(ValueError, match='Invalid number of qubits'):
        transforms.bravyi_kitaev(ops.FermionOperator('4^ 4'), 3)
    with pytest.raises(TypeError):
        transforms.jordan_wigner(ops.QubitOperator('X0'))
