"""CPU: pins the oracle (oracle/pauli_oracle.py, oracle/pauli_oracle.c) against the
reference's known-answer tests and against outputs of the real reference frozen in
tests/golden/ by oracle/make_golden.py."""
import numpy
import pytest

import c_oracle
import pauli_oracle as po
import helpers
import known_answers as ka
from openfermion_b200 import compiler, hamiltonians
from openfermion_b200.ops import QubitOperator, count_qubits


@pytest.mark.parametrize('name,op,n,vec,expected', ka.MATVEC, ids=[k[0] for k in ka.MATVEC])
def test_matvec_known_answers(name, op, n, vec, expected):
    n = count_qubits(op) if n is None else n
    assert numpy.allclose(po.matvec(op.terms, n, vec), expected)


def test_matvec_compare_with_sparse():
    # linear_qubit_operator_test.py:178-188
    op = QubitOperator('X0 Y1 Z3')
    mat = po.qubit_operator_sparse(op.terms).toarray()
    cols = numpy.array([po.matvec(op.terms, 4, v) for v in numpy.identity(16)]).T
    assert numpy.allclose(cols, mat)


@pytest.mark.parametrize('name,op,n,data,indices,shape', ka.CSC, ids=[k[0] for k in ka.CSC])
def test_csc_known_answers(name, op, n, data, indices, shape):
    mat = po.qubit_operator_sparse(op.terms, n)
    assert mat.shape == shape
    assert list(mat.data) == data
    assert list(mat.indices) == indices


@pytest.mark.parametrize('name,op,n,expected', ka.DIAGONAL, ids=[k[0] for k in ka.DIAGONAL])
def test_diagonal_known_answers(name, op, n, expected):
    n = count_qubits(op) if n is None else n
    assert numpy.allclose(po.diagonal(op.terms, n), expected)


@pytest.mark.parametrize('string', ['Z1 X2 Y5', 'Z1 Z2 Z5'])
def test_diagonal_equals_matvec_on_identity(string):
    # sparse_tools_test.py:225-244
    op = QubitOperator(string)
    want = numpy.diag(po.matmat(op.terms, 6, numpy.eye(2**6)))
    assert numpy.allclose(po.diagonal(op.terms, 6), want)


def test_diagonal_complex_coefficient_raises_like_reference():
    op = QubitOperator('Z0', 1.0 + 0.0j)
    with pytest.raises(TypeError):
        po.diagonal(op.terms, 1)


def test_invalid_qubit_count():
    with pytest.raises(ValueError):
        po.qubit_operator_sparse(QubitOperator('X3').terms, 1)


@pytest.mark.parametrize('case', helpers.qubit_cases(), ids=lambda c: c.name)
def test_oracle_matches_reference_outputs(case):
    op = helpers.qubit_case_operator(case)
    n = case.n
    assert numpy.array_equal(po.matvec(op.terms, n, case['vec']), case['matvec'])
    mat = po.qubit_operator_sparse(op.terms, n)
    helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], rtol=0)
    assert abs(po.expectation_terms(op.terms, n, case['vec']) - complex(case['expectation'])) \
        <= 1e-13 * abs(complex(case['expectation'])) + 1e-300
    if case.has('diagonal'):
        assert numpy.array_equal(po.diagonal(op.terms, n), case['diagonal'])


@pytest.mark.parametrize('case', helpers.fermion_cases(), ids=lambda c: c.name)
def test_oracle_fermion_route_matches_reference(case):
    op = helpers.fermion_operator_from(case['terms'], case['coeffs'])
    mat = po.jordan_wigner_sparse(op.terms, case.n)
    helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], rtol=0)


def test_oracle_ladder_operator_matches_reference():
    # jordan_wigner_ladder_sparse (sparse_tools.py:49-73) = the one-factor fermion route
    data = helpers.golden('ladder_cases.npz')
    for name in data['names']:
        n, mode, kind = (int(v) for v in data[name + '/args'])
        mat = po.jordan_wigner_sparse({((mode, kind),): 1.0}, n)
        helpers.assert_csc_matches(mat, data[name + '/indptr'], data[name + '/indices'],
                                   data[name + '/data'], rtol=0)


def test_kronecker_helpers_match_reference():
    # wrapped_kronecker / kronecker_operators (sparse_tools.py:39-46): generic host helpers
    import scipy.sparse
    from openfermion_b200.linalg import kronecker_operators, wrapped_kronecker
    data = helpers.golden('ladder_cases.npz')
    a, b, c = (scipy.sparse.csc_matrix(data['kron_in_' + t]) for t in 'abc')
    for tag, mat in (('ab', wrapped_kronecker(a, b)), ('abc', kronecker_operators([a, b, c]))):
        assert mat.format == 'csc' and tuple(data['kron_%s/shape' % tag]) == mat.shape
        mat.sort_indices()
        want = scipy.sparse.csc_matrix((data['kron_%s/data' % tag], data['kron_%s/indices' % tag],
                                        data['kron_%s/indptr' % tag]), shape=mat.shape)
        assert abs(mat - want).max() == 0


@pytest.mark.parametrize('case', helpers.qubit_cases(), ids=lambda c: c.name)
def test_c_oracle_matches_numpy_oracle(case):
    op = helpers.qubit_case_operator(case)
    x, z, c = compiler.compile_qubit_terms(op.terms, case.n)
    got = c_oracle.matvec(case.n, x, z, c, case['vec'])
    assert helpers.close(got, case['matvec'], 1e-15)
    if case.has('diagonal'):
        assert helpers.close(c_oracle.diagonal(case.n, x, z, c), case['diagonal'], 1e-15)


@pytest.mark.parametrize('case', helpers.large_cases(), ids=lambda c: c.name)
def test_c_oracle_matches_reference_at_tiled_kernel_sizes(case):
    """SURVEY.md section 8d parity ladder: sampled outputs of the real reference's matvec at
    14-24 qubits (tests/golden/large_cases.npz) pin the C oracle at the sizes the tiled CUDA
    kernel runs at; the GPU suite checks the kernel against the same samples."""
    import hashlib
    n = case.n
    v = helpers.large_state(n, int(case['seed']))
    assert hashlib.sha256(numpy.ascontiguousarray(v).tobytes()).hexdigest() == str(case['vec_sha256'])
    op = helpers.qubit_case_operator(case)
    x, z, c = compiler.compile_qubit_terms(op.terms, n)
    got = c_oracle.matvec(n, x, z, c, v)
    scale = float(case['max_abs'])
    sample = case['sample_index']
    assert numpy.max(numpy.abs(got[sample] - case['sample_matvec'])) <= 1e-13 * scale
    one_norm = float(numpy.sum(numpy.abs(case['coeffs'])))
    assert abs(numpy.vdot(v, got) - complex(case['expectation'])) <= 1e-13 * one_norm
    assert abs(numpy.linalg.norm(got) - float(case['norm'])) <= 1e-12 * float(case['norm'])
    if case.has('sample_diagonal'):
        diag = c_oracle.diagonal(n, x, z, c)
        assert numpy.max(numpy.abs(diag[sample] - case['sample_diagonal'])) <= 1e-13 * numpy.max(numpy.abs(diag))


def _csc_ladder_cases():
    data = helpers.golden('csc_ladder_cases.npz')
    return [helpers.Case(data, str(n)) for n in data['names']]


@pytest.mark.parametrize('case', [c for c in _csc_ladder_cases() if c.n <= 12 or 'hubbard' in c.name],
                         ids=lambda c: c.name)
def test_oracle_csc_matches_reference_hashes(case):
    """CSC ladder (tests/golden/csc_ladder_cases.npz): the scipy-based oracle against sha256 of the
    real reference's matrices.  The 14-qubit molecular cases need gigabytes of triplets, as in the
    reference, and are left to the GPU suite."""
    import hashlib

    def sha(a):
        return hashlib.sha256(numpy.ascontiguousarray(a).tobytes()).hexdigest()
    if str(case['route']) == 'pauli':
        mat = po.qubit_operator_sparse(helpers.qubit_case_operator(case).terms, case.n)
    else:
        op = helpers.fermion_operator_from(case['terms'], case['coeffs'], case['coeff_is_complex'])
        mat = po.jordan_wigner_sparse(op.terms, case.n)
    assert mat.nnz == int(case['nnz']) and str(mat.indices.dtype) == str(case['index_dtype'])
    assert sha(mat.indptr) == str(case['indptr_sha256'])
    assert sha(mat.indices) == str(case['indices_sha256'])
    assert sha(mat.data) == str(case['data_sha256'])


def test_ground_state_known_answer():
    # sparse_tools_test.py:593-604
    op = QubitOperator('Y0 X1') + QubitOperator('Z0 Z1')
    energy, state = po.get_ground_state(po.qubit_operator_sparse(op.terms))
    expected = numpy.array([0, 1j, 1, 0]) / numpy.sqrt(2.0)
    assert abs(energy + 2) < 1e-7
    assert abs(abs(numpy.vdot(expected, state)) - 1.0) < 1e-7


def test_lih_ground_state_matches_file_energy():
    # testing/lih_integration_test.py:70-71,103-106
    data = helpers.golden('lih_sto3g.npz')
    ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                       data['two_body_integrals'])
    from openfermion_b200.linalg.sparse_tools import _tensor_fermion_terms
    fop = _tensor_fermion_terms(ham)
    mat = po.jordan_wigner_sparse(fop.terms, 12)
    assert mat.nnz == int(data['nnz'])
    assert numpy.array_equal(mat.indptr, data['indptr'])
    import hashlib
    assert hashlib.sha256(mat.indices.tobytes()).hexdigest() == str(data['indices_sha256'])
    energy, _ = po.get_ground_state(mat)
    assert abs(energy - float(data['fci_energy'])) < 1e-7


def test_oracle_davidson_reference_trajectory():
    # davidson_test.py:205-215: exactly two iterations from e_0 give 1.41556103, not converged
    dimension = 10
    numpy.random.seed(dimension)
    rand = numpy.array(numpy.random.rand(dimension, dimension))
    numpy.random.seed(dimension * 2)
    diag = numpy.array(range(dimension)) + numpy.random.rand(dimension)
    matrix = rand + rand.conj().T + numpy.diag(diag)
    guess = numpy.eye(dimension, 10)
    ok, vals, _ = po.davidson_lowest_n(lambda v: matrix @ v, numpy.diag(matrix), 1, guess[:, :1],
                                       max_iterations=2)
    assert not ok and numpy.allclose(vals, [1.41556103])
    ok, vals, _ = po.davidson_lowest_n(lambda v: matrix @ v, numpy.diag(matrix), 2, guess[:, :2],
                                       max_iterations=5)
    assert ok and numpy.allclose(vals, [1.15675714, 1.59132505])
    ok, vals, _ = po.davidson_lowest_n(lambda v: matrix @ v, numpy.diag(matrix), 2, guess[:, :2],
                                       max_iterations=5, max_subspace=8)
    assert not ok and numpy.allclose(vals, [1.1572995, 1.61393264])


@pytest.mark.needs_reference
def test_oracle_against_live_reference():
    """Build container only: fresh random operators through the real reference."""
    import ref_shim
    ref_shim.load()
    from openfermion.linalg import LinearQubitOperator, qubit_operator_sparse
    from openfermion.testing.testing_utils import random_qubit_operator
    rng = numpy.random.default_rng(5)
    for seed in range(100, 106):
        n = 3 + seed % 6
        op = random_qubit_operator(n, 20, n, seed=seed)
        v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
        assert numpy.array_equal(LinearQubitOperator(op, n) * v, po.matvec(op.terms, n, v))
        ref = qubit_operator_sparse(op, n)
        helpers.assert_csc_matches(po.qubit_operator_sparse(op.terms, n), ref.indptr, ref.indices,
                                   ref.data, rtol=0)
