"""GPU: parity of the CUDA path (through the Python mirror -> C ABI) against the oracle,
the reference's known answers and the golden vectors generated from the real reference.
Bars (BASELINE.json): CSC indptr/indices bit-exact; values, matvecs, expectations within
1e-12 relative."""
import hashlib

import numpy
import pytest
import scipy.sparse

import c_oracle
import helpers
import known_answers as ka
import pauli_oracle as po
from openfermion_b200 import compiler, hamiltonians
from openfermion_b200.linalg import (LinearQubitOperator, ParallelLinearQubitOperator,
                                     LinearQubitOperatorOptions, expectation, variance,
                                     get_linear_qubit_operator_diagonal, get_sparse_operator,
                                     jordan_wigner_sparse, qubit_operator_sparse)
from openfermion_b200.ops import QubitOperator, count_qubits

pytestmark = pytest.mark.gpu
RTOL = 1e-12
# Golden Pauli-route cases with more than 16 triplets per column and generic real
# coefficients: the reference's structure contains |v| ~ 1e-16 cancellation residues whose
# survival depends on the internal order of scipy's unstable per-column std::sort
# (SURVEY.md section 8c).  The default ('auto') CSC mode replays that order and is compared
# bit-exactly; the fast mode is compared modulo those residues.
RESIDUE_CASES = {'molecular_6_jw', 'molecular_6_bk', 'molecular_8_jw'}


# ---- reference known answers -----------------------------------------------------
@pytest.mark.parametrize('name,op,n,vec,expected', ka.MATVEC, ids=[k[0] for k in ka.MATVEC])
def test_matvec_known_answers(name, op, n, vec, expected):
    got = LinearQubitOperator(op, n) * vec
    assert got.dtype == numpy.complex128 and got.shape == vec.shape
    assert numpy.allclose(got, expected, rtol=0, atol=1e-14)


def test_matvec_wrong_length_raises():
    with pytest.raises(ValueError):
        LinearQubitOperator(QubitOperator('X3')) * numpy.zeros(4)


def test_matvec_compare_with_sparse():
    op = QubitOperator('X0 Y1 Z3')
    mat = qubit_operator_sparse(op).toarray()
    lin = LinearQubitOperator(op)
    cols = numpy.array([lin * v for v in numpy.identity(16)]).T
    assert numpy.array_equal(cols, mat)
    assert numpy.array_equal(lin.matmat(numpy.identity(16, dtype=complex)), mat)


@pytest.mark.parametrize('name,op,n,data,indices,shape', ka.CSC, ids=[k[0] for k in ka.CSC])
def test_csc_known_answers(name, op, n, data, indices, shape):
    mat = get_sparse_operator(op, n)
    assert scipy.sparse.isspmatrix_csc(mat) and mat.shape == shape
    assert list(mat.data) == data
    assert list(mat.indices) == indices
    ref = po.qubit_operator_sparse(op.terms, n)
    assert mat.dtype == ref.dtype and mat.indices.dtype == ref.indices.dtype
    assert numpy.array_equal(mat.indptr, ref.indptr)


@pytest.mark.parametrize('name,op,n,expected', ka.DIAGONAL, ids=[k[0] for k in ka.DIAGONAL])
def test_diagonal_known_answers(name, op, n, expected):
    got = get_linear_qubit_operator_diagonal(op, n)
    assert got.dtype == numpy.float64
    assert numpy.array_equal(got, expected)


@pytest.mark.parametrize('string', ['Z1 X2 Y5', 'Z1 Z2 Z5'])
def test_diagonal_equals_operator_on_identity(string):
    op = QubitOperator(string)
    want = numpy.diag(LinearQubitOperator(op).matmat(numpy.eye(2**6, dtype=complex)))
    assert numpy.allclose(get_linear_qubit_operator_diagonal(op), want)


def test_diagonal_complex_coefficient_raises():
    with pytest.raises(TypeError):
        get_linear_qubit_operator_diagonal(QubitOperator('Z0', 1.0 + 0j))


def test_expectation_known_answers():
    # sparse_tools_test.py:607-637
    vector = numpy.array([0.0, 1.0j, 0.0, 1.0j])
    sparse = get_sparse_operator(QubitOperator('X0'), n_qubits=2)
    assert abs(expectation(sparse, vector) - 2.0) < 1e-14
    density = scipy.sparse.csc_matrix(numpy.outer(vector, numpy.conjugate(vector)))
    assert abs(expectation(sparse, density) - 2.0) < 1e-14
    lin = LinearQubitOperator(QubitOperator('X0'), n_qubits=2)
    assert abs(expectation(lin, vector) - 2.0) < 1e-14
    assert abs(expectation(sparse, vector.reshape(4, 1)) - 2.0) < 1e-14
    assert abs(expectation(sparse, numpy.array([1j, -1j, -1j, -1j]))) < 1e-14
    with pytest.raises(ValueError):
        expectation(lin, scipy.sparse.csc_matrix(numpy.array([1, 0, 0, 0])))
    with pytest.raises(ValueError):
        expectation(scipy.sparse.csc_matrix(numpy.array([1, 0, 0, 0])), lin)


def test_variance_known_answers():
    # sparse_tools_test.py:640-671
    X = get_sparse_operator(QubitOperator('X0'))
    Z = get_sparse_operator(QubitOperator('Z0'))
    zero, plus = numpy.array([1.0, 0.0]), numpy.array([1.0, 1.0]) / numpy.sqrt(2)
    assert abs(variance(X, zero) - 1.0) < 1e-13 and abs(variance(Z, zero)) < 1e-13
    assert abs(variance(Z, plus) - 1.0) < 1e-13 and abs(variance(X, plus)) < 1e-13
    lin = LinearQubitOperator(QubitOperator('X0') + QubitOperator('Z0', 0.5))
    mat = get_sparse_operator(QubitOperator('X0') + QubitOperator('Z0', 0.5))
    v = numpy.array([0.6, 0.8j])
    assert abs(variance(lin, v) - variance(mat, v)) < 1e-13


def test_variance_matrix_free_on_the_device_equals_the_oracle():
    # the device-resident <v|H(Hv)> - <v|Hv>^2 of the matrix-free operators against the formula of
    # sparse_tools.py:674-690 evaluated with the oracle's matvec; complex, non-Hermitian sum too
    rng = numpy.random.default_rng(8)
    for op, n in ((hamiltonians.molecular_jw(7, seed=5), 14), (hamiltonians.random_pauli_sum(11, 90, 3, True), 11)):
        v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        v /= numpy.linalg.norm(v)
        w = po.matvec(op.terms, n, v)
        want = numpy.vdot(v, po.matvec(op.terms, n, w)) - numpy.vdot(v, w) ** 2
        got = variance(LinearQubitOperator(op, n), v)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
    with pytest.raises(ValueError):
        variance(LinearQubitOperator(hamiltonians.hubbard_jw(2, 2), 8), numpy.ones(17))


# ---- golden vectors from the real reference ------------------------------------------
@pytest.mark.parametrize('case', helpers.qubit_cases(), ids=lambda c: c.name)
def test_golden_matvec_expectation_diagonal(case):
    op = helpers.qubit_case_operator(case)
    lin = LinearQubitOperator(op, case.n)
    assert helpers.close(lin * case['vec'], case['matvec'], RTOL)
    want = complex(case['expectation'])
    scale = numpy.sum(numpy.abs(case['coeffs'])) * numpy.vdot(case['vec'], case['vec']).real
    assert abs(expectation(lin, case['vec']) - want) <= RTOL * scale
    if case.has('diagonal'):
        assert helpers.close(get_linear_qubit_operator_diagonal(op, case.n), case['diagonal'], RTOL)


@pytest.mark.parametrize('case', helpers.qubit_cases(), ids=lambda c: c.name)
def test_golden_csc_pauli_route(case):
    """Indices AND values bit-exact against the reference's own output for every golden case
    (values compared with ==; the only tolerated difference is the sign of a zero part)."""
    op = helpers.qubit_case_operator(case)
    mat = qubit_operator_sparse(op, case.n)
    helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], rtol=0)


@pytest.mark.parametrize('name', sorted(RESIDUE_CASES))
def test_golden_csc_pauli_route_fast_mode_contract(name, monkeypatch):
    monkeypatch.setenv('OFB_CSC_MODE', 'fast')
    case = helpers.Case(helpers.golden('qubit_cases.npz'), name)
    mat = qubit_operator_sparse(helpers.qubit_case_operator(case), case.n)
    helpers.assert_csc_matches_modulo_residues(
        mat, case['indptr'], case['indices'], case['data'],
        float(numpy.sum(numpy.abs(case['coeffs']))), RTOL)


def test_csc_exact_mode_generic_molecular_n10_bit_exact():
    """Generic (non-dyadic) coefficients, 1 519 terms per column: the scipy-order mode must
    reproduce the oracle's (= scipy's) structure and values bit for bit."""
    op = hamiltonians.molecular_jw(5, seed=1234)
    got = qubit_operator_sparse(op, 10)
    want = po.qubit_operator_sparse(op.terms, 10)
    helpers.assert_csc_matches(got, want.indptr, want.indices, want.data, rtol=0)


def test_csc_exact_mode_multi_batch_and_fermion_route(monkeypatch):
    monkeypatch.setenv('OFB_CSC_MODE', 'exact')
    monkeypatch.setenv('OFB_SORT_BATCH', '256')  # several column batches, re-sorted in the fill pass
    data = helpers.golden('fermion_cases.npz')
    for name in ('molecular_8_dense', 'hubbard_2x3_open', 'molecular_4_complex'):
        case = helpers.Case(data, name)
        fop = helpers.fermion_operator_from(case['terms'], case['coeffs'])
        mat = jordan_wigner_sparse(fop, case.n)
        helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], rtol=0)
    case = helpers.Case(helpers.golden('qubit_cases.npz'), 'molecular_8_jw')
    mat = qubit_operator_sparse(helpers.qubit_case_operator(case), case.n)
    helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], rtol=0)


@pytest.mark.parametrize('case', helpers.fermion_cases(), ids=lambda c: c.name)
def test_golden_csc_fermion_route(case):
    op = helpers.fermion_operator_from(case['terms'], case['coeffs'])
    mat = jordan_wigner_sparse(op, case.n)
    helpers.assert_csc_matches(mat, case['indptr'], case['indices'], case['data'], RTOL)
    assert mat.has_canonical_format


def test_lih_sparse_operator_matches_reference(monkeypatch):
    data = helpers.golden('lih_sto3g.npz')
    ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                       data['two_body_integrals'])
    mat = get_sparse_operator(ham)  # rows built array-wise from the tensors
    monkeypatch.setenv('OFB_TENSOR_ROWS', '0')
    via_terms = get_sparse_operator(ham)  # through the FermionOperator, as the reference converts
    monkeypatch.delenv('OFB_TENSOR_ROWS')
    assert numpy.array_equal(mat.indptr, via_terms.indptr) and numpy.array_equal(mat.indices, via_terms.indices)
    assert mat.data.tobytes() == via_terms.data.tobytes()
    assert mat.nnz == int(data['nnz']) and str(mat.indices.dtype) == str(data['indices_dtype'])
    assert numpy.array_equal(mat.indptr, data['indptr'])
    assert hashlib.sha256(mat.indices.tobytes()).hexdigest() == str(data['indices_sha256'])
    assert numpy.array_equal(mat.indices[:512], data['indices_head'])
    assert helpers.close(mat.data[:512], data['data_head'], RTOL)
    assert abs(numpy.abs(mat.data).sum() - float(data['data_abs_sum'])) <= 1e-12 * float(data['data_abs_sum'])
    hf_state = numpy.zeros(2**12)
    hf_state[sum(2 ** (11 - i) for i in range(4))] = 1.0
    assert abs(expectation(mat, hf_state) - complex(data['reference_hf_expectation'])) < 1e-12
    qop = helpers.qubit_operator_from(data['qubit_terms'], data['qubit_coeffs'])
    assert abs(expectation(LinearQubitOperator(qop, 12), hf_state + 0j)
               - complex(data['reference_hf_expectation'])) < 1e-11


# ---- larger sizes against the C oracle ---------------------------------------------------
def _state(n, seed):
    rng = numpy.random.default_rng(seed)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / numpy.linalg.norm(v)


@pytest.mark.parametrize('name,builder', [
    ('hubbard_2x4_n16', lambda: hamiltonians.hubbard_jw(2, 4)),
    ('hubbard_2x5_n20', lambda: hamiltonians.hubbard_jw(2, 5)),
    ('molecular_n12', lambda: hamiltonians.molecular_jw(6, seed=1234)),
    ('molecular_n16', lambda: hamiltonians.molecular_jw(8, seed=1234)),
    ('dense_paulis_complex_n15', lambda: hamiltonians.random_pauli_sum(15, 300, 3, True)),
    ('dense_paulis_real_n13', lambda: hamiltonians.random_pauli_sum(13, 500, 4, False)),
    ('hubbard_3x4_n24', lambda: hamiltonians.hubbard_jw(3, 4)),
    ('molecular_bk_n16', lambda: hamiltonians.molecular_jw(8, seed=1234, encoding='bk')),
    ('hubbard_2x5_bk_n20', lambda: hamiltonians.hubbard_jw(2, 5, encoding='bk')),
])
def test_matvec_and_expectation_vs_c_oracle(name, builder):
    op = builder()
    n = op.n_qubits
    # the masks handed to the C oracle come from the ORACLE's own fold of the term dict
    # (pauli_oracle.term_masks, linear_qubit_operator.py:122-134), not from the product compiler,
    # so a compiler bug cannot cancel out; the compiler's tables are compared too
    x = numpy.zeros(len(op.terms), dtype=numpy.uint64)
    z = numpy.zeros(len(op.terms), dtype=numpy.uint64)
    c = numpy.zeros(len(op.terms), dtype=numpy.complex128)
    for k, (term, coefficient) in enumerate(op.terms.items()):
        x[k], z[k], _ = po.term_masks(term, n)
        c[k] = coefficient
    cx, cz, cc = compiler.compile_qubit_operator(op, n)
    assert sorted(zip(x.tolist(), z.tolist(), c.tolist())) == sorted(zip(cx.tolist(), cz.tolist(), cc.tolist()))
    v = _state(n, 7)
    want = c_oracle.matvec(n, x, z, c, v)
    lin = LinearQubitOperator(op, n)
    got = lin * v
    assert helpers.close(got, want, RTOL)
    e_want = numpy.vdot(v, want)
    assert abs(expectation(lin, v) - e_want) <= RTOL * numpy.sum(numpy.abs(c))
    if not numpy.any(c.imag):
        assert helpers.close(get_linear_qubit_operator_diagonal(op, n), c_oracle.diagonal(n, x, z, c), RTOL)


@pytest.mark.parametrize('case', helpers.large_cases(), ids=lambda c: c.name)
def test_matvec_matches_reference_samples_at_tiled_kernel_sizes(case):
    """Sampled outputs of the REAL reference's LinearQubitOperator._matvec at 14-24 qubits
    (tests/golden/large_cases.npz, SURVEY.md section 8d parity ladder): the tiled kernel against
    the reference itself, 1e-12 relative to max |H v|; expectation and diagonal likewise."""
    n = case.n
    v = helpers.large_state(n, int(case['seed']))
    op = helpers.qubit_case_operator(case)
    lin = LinearQubitOperator(op, n)
    got = lin * v
    scale = float(case['max_abs'])
    sample = case['sample_index']
    assert numpy.max(numpy.abs(got[sample] - case['sample_matvec'])) <= RTOL * scale
    assert abs(numpy.linalg.norm(got) - float(case['norm'])) <= RTOL * float(case['norm'])
    assert abs(numpy.sum(got) - complex(case['sum'])) <= 1e-10 * float(case['norm']) * 2.0 ** (n / 2)
    one_norm = float(numpy.sum(numpy.abs(case['coeffs'])))  # bound of |<v|H|v>| for normalised v
    assert abs(expectation(lin, v) - complex(case['expectation'])) <= RTOL * one_norm
    if case.has('sample_diagonal'):
        diag = get_linear_qubit_operator_diagonal(op, n)
        assert diag.dtype == numpy.float64
        assert numpy.max(numpy.abs(diag[sample] - case['sample_diagonal'])) <= RTOL * numpy.max(numpy.abs(diag))
        assert abs(numpy.sum(diag) - float(case['diagonal_sum'])) <= 1e-10 * numpy.sum(numpy.abs(diag))


def test_matvec_accepts_reference_input_kinds():
    op = hamiltonians.hubbard_jw(2, 2)
    lin = LinearQubitOperator(op, 8)
    ints = numpy.arange(256)
    want = po.matvec(op.terms, 8, ints)
    assert helpers.close(lin * ints, want, RTOL)
    col = (ints * 0.5).reshape(256, 1)
    out = lin * col
    assert out.shape == (256, 1) and helpers.close(out[:, 0], want * 0.5, RTOL)
    X = numpy.stack([_state(8, s) for s in range(5)], axis=1)
    assert helpers.close(lin.matmat(X), po.matmat(op.terms, 8, X), RTOL)
    assert helpers.close(lin @ X, po.matmat(op.terms, 8, X), RTOL)
    with pytest.raises(NotImplementedError):
        lin.rmatvec(ints)


def test_padded_qubits_and_single_qubit_and_zero_operator():
    op = QubitOperator('X0 Y1', 0.5) + QubitOperator('Z1', -1.25)
    for n in (2, 5, 11, 15):
        v = _state(n, n)
        assert helpers.close(LinearQubitOperator(op, n) * v, po.matvec(op.terms, n, v), RTOL)
    one = QubitOperator('Y0', 2.0)
    assert helpers.close(LinearQubitOperator(one) * numpy.array([1.0, 2.0]), [-4j, 2j], RTOL)
    assert numpy.array_equal(LinearQubitOperator(QubitOperator(), 3) * numpy.ones(8), numpy.zeros(8))
    ident = LinearQubitOperator(QubitOperator((), 3.0), 0)
    assert numpy.array_equal(ident * numpy.array([2.0]), [6.0])


def test_parallel_operator_matches_fused():
    # linear_qubit_operator_test.py:191-284 semantics on one GPU
    op = hamiltonians.hubbard_jw(2, 3)
    qop = QubitOperator()
    qop.terms = dict(op.terms)
    par = ParallelLinearQubitOperator(qop, 12, LinearQubitOperatorOptions(processes=3))
    v = _state(12, 1)
    want = po.matvec(qop.terms, 12, v)
    assert helpers.close(par * v, want, RTOL)
    parts = sum(lin * v for lin in par.linear_operators)
    assert helpers.close(parts, want, RTOL)
    assert par.options.pool is None
    empty = ParallelLinearQubitOperator(QubitOperator(), 2)
    assert numpy.array_equal(empty * numpy.ones(4), numpy.zeros(4))


# ---- size-independent properties at sizes the oracle cannot reach ---------------------------
@pytest.mark.parametrize('builder,n', [(lambda: hamiltonians.hubbard_jw(3, 4), 24),
                                       (lambda: hamiltonians.molecular_jw(11, seed=1234), 22)])
def test_hermiticity_and_linearity_properties(builder, n):
    lin = LinearQubitOperator(builder(), n)
    u, v = _state(n, 1), _state(n, 2)
    hu, hv = lin * u, lin * v
    # <u|Hv> = conj(<v|Hu>) for a Hermitian H
    assert abs(numpy.vdot(u, hv) - numpy.conj(numpy.vdot(v, hu))) < 1e-11
    comb = lin * (0.3 * u - 1.7j * v)
    assert helpers.close(comb, 0.3 * hu - 1.7j * hv, 1e-12)
    assert abs(expectation(lin, u) - numpy.vdot(u, hu)) < 1e-11
    assert abs(expectation(lin, u).imag) < 1e-11


# ---- CSC at larger sizes ----------------------------------------------------------------------
def test_csc_dyadic_molecular_bit_exact():
    """Dyadic coefficients: every column sum is exact, so structure and values are bit-exact
    under any summation order (SURVEY.md section 8c finding 3)."""
    op = hamiltonians.molecular_jw(5, seed=1234, dyadic_bits=16)
    got = qubit_operator_sparse(op, 10)
    want = po.qubit_operator_sparse(op.terms, 10)
    helpers.assert_csc_matches(got, want.indptr, want.indices, want.data, rtol=0)


def test_csc_hubbard_n12_and_n16_bit_exact():
    for dims in ((2, 3), (2, 4)):
        op = hamiltonians.hubbard_jw(*dims)
        got = qubit_operator_sparse(op, op.n_qubits)
        want = po.qubit_operator_sparse(op.terms, op.n_qubits)
        helpers.assert_csc_matches(got, want.indptr, want.indices, want.data, rtol=0)


def test_csc_generic_coefficients_residue_tolerant():
    """Generic real coefficients: scipy's per-column std::sort (unstable above 16 entries)
    decides which cancellation residues survive, so structure is compared after dropping
    entries below 64*eps*sum|c_t| on both sides; all other entries must agree to 1e-12."""
    op = hamiltonians.molecular_jw(5, seed=1234)
    got = qubit_operator_sparse(op, 10).tocoo()
    want = po.qubit_operator_sparse(op.terms, 10).tocoo()
    floor = 64 * numpy.finfo(float).eps * numpy.sum(numpy.abs(op.coeff))

    def strong(m):
        keep = numpy.abs(m.data) > floor
        return scipy.sparse.csc_matrix((m.data[keep], (m.row[keep], m.col[keep])), shape=m.shape)

    a, b = strong(got), strong(want)
    assert numpy.array_equal(a.indptr, b.indptr) and numpy.array_equal(a.indices, b.indices)
    assert numpy.max(numpy.abs(a.data - b.data)) <= 1e-12 * numpy.max(numpy.abs(b.data))
    dense_diff = abs(got.tocsc() - want.tocsc())
    assert dense_diff.max() <= floor


def test_csc_matches_matrix_free_operator_columns():
    op = hamiltonians.random_pauli_sum(7, 60, 9, True)
    mat = qubit_operator_sparse(op, 7).toarray()
    cols = LinearQubitOperator(op, 7).matmat(numpy.eye(128, dtype=complex))
    assert helpers.close(mat, cols, RTOL)


def test_csc_fermion_route_equals_pauli_route_spectrum():
    # sparse_tools_test.py:126-138
    data = helpers.golden('fermion_cases.npz')
    case = helpers.Case(data, 'molecular_6')
    fop = helpers.fermion_operator_from(case['terms'], case['coeffs'])
    f = jordan_wigner_sparse(fop, 6).toarray()
    q = qubit_operator_sparse(hamiltonians.molecular_jw(3, seed=5), 6).toarray()
    assert numpy.allclose(numpy.linalg.eigvalsh(f), numpy.linalg.eigvalsh(q), atol=1e-10)


def test_ladder_operator_matrices_match_reference():
    # jordan_wigner_ladder_sparse (sparse_tools.py:49-73), golden matrices from the reference
    from openfermion_b200.linalg import jordan_wigner_ladder_sparse
    data = helpers.golden('ladder_cases.npz')
    for name in data['names']:
        n, mode, kind = (int(v) for v in data[name + '/args'])
        mat = jordan_wigner_ladder_sparse(n, mode, kind)
        assert mat.shape == (1 << n, 1 << n) and mat.dtype == numpy.complex128
        helpers.assert_csc_matches(mat, data[name + '/indptr'], data[name + '/indices'],
                                   data[name + '/data'], rtol=0)
    with pytest.raises(ValueError):
        jordan_wigner_ladder_sparse(3, 3, 1)


# ---- sharded kernel form on one GPU (virtual ranks) -------------------------------------------
@pytest.mark.parametrize('basis_kind', ['natural', 'symmetry', 'symmetry+locality'])
@pytest.mark.parametrize('builder,n,world', [(lambda: hamiltonians.hubbard_jw(2, 4), 16, 4),
                                             (lambda: hamiltonians.hubbard_jw(2, 4), 16, 8),
                                             (lambda: hamiltonians.molecular_jw(6, seed=1234), 12, 8),
                                             (lambda: hamiltonians.molecular_jw(5, seed=3, encoding='bk'), 10, 4),
                                             (lambda: hamiltonians.random_pauli_sum(13, 200, 2, True), 13, 2)])
def test_virtual_rank_sharding_matches_unsharded(builder, n, world, basis_kind):
    """Every rank's shard is another region of one device buffer (the p2p data path with
    pointers into the same GPU): ofb_apply_groups with local_bits < n, a per-rank index_base,
    per-pattern group ranges and accumulate must reproduce the unsharded H|psi> -- with the
    shards cut by the top qubits and by conserved parities (shard_basis.py: relabelled masks,
    relabelled state, same kernel)."""
    from openfermion_b200 import _lib
    from openfermion_b200.distributed import ShardLayout
    from openfermion_b200.linalg import krylov
    from openfermion_b200.plan import PauliPlan
    from openfermion_b200.shard_basis import ShardBasis
    op = builder()
    lin = LinearQubitOperator(op, n)
    x, z, c = compiler.compile_qubit_operator(op, n)
    if basis_kind == 'natural':
        basis, plan = ShardBasis.natural(n), lin.plan
    else:
        basis = ShardBasis.for_sharding(n, world.bit_length() - 1, x,
                                        local_order='locality' if 'locality' in basis_kind else 'natural')
        x, z, c = basis.transform_terms(x, z, c)
        plan = PauliPlan(n, x, z, c)
    layout = ShardLayout(n, world, x)
    assert numpy.array_equal(layout.group_x, plan.group_x_masks())
    v = _state(n, 11)
    want = lin * v
    d_in = _lib.to_device(basis.from_natural(v))
    d_out = _lib.DeviceBuffer(v.nbytes)
    ld = layout.local_dim
    for rank in range(world):
        first = True
        for x_hi, g0, g1 in layout.patterns:
            src = d_in.ptr + layout.peer(rank, x_hi) * ld * 16
            plan.apply_groups_device(src, d_out.ptr + rank * ld * 16, layout.local_bits,
                                     layout.index_base(rank), g0, g1, 0 if first else 1)
            first = False
    got = basis.to_natural(krylov.download(1 << n, d_out.ptr))
    assert helpers.close(got, want, 1e-13)


@pytest.mark.parametrize('n,local_bits', [(36, 11), (36, 7), (33, 12), (40, 10)])
def test_shards_of_states_beyond_32_qubits(n, local_bits):
    """More than 32 qubits only exist sharded; the kernel's 64-bit sign path (z bits above
    bit 31 against the rank's index_base) is checked on single shards against the gather
    formula evaluated with Python integers."""
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import krylov
    from openfermion_b200.plan import PauliPlan
    rng = numpy.random.default_rng(n * 100 + local_bits)
    T = 60
    low = (1 << local_bits) - 1
    # all x masks share their high part, so one peer shard feeds every term
    x_hi = int(rng.integers(1, 1 << 20)) << local_bits
    xs = [x_hi | int(rng.integers(0, 1 << local_bits)) for _ in range(T)]
    xs[:10] = [x_hi | (xs[0] & low)] * 10  # a big group
    zs = [int(rng.integers(0, 1 << 62)) & ((1 << n) - 1) for _ in range(T)]
    cs = rng.normal(size=T) + 1j * rng.normal(size=T) * (rng.random(T) < 0.3)
    plan = PauliPlan(n, numpy.array(xs, dtype=numpy.uint64), numpy.array(zs, dtype=numpy.uint64), cs)
    ld = 1 << local_bits
    rank = int(rng.integers(0, 1 << (n - local_bits)))
    base = rank << local_bits
    vin = rng.normal(size=ld) + 1j * rng.normal(size=ld)
    want = numpy.zeros(ld, dtype=complex)
    for x, z, c in zip(xs, zs, cs):
        phase = c * (-1j) ** bin(x & z).count('1')
        for j in range(ld):
            sign = -1.0 if bin((base | j) & z).count('1') & 1 else 1.0
            want[j] += phase * sign * vin[j ^ (x & low)]
    d_in = _lib.to_device(vin)
    d_out = _lib.DeviceBuffer(vin.nbytes)
    plan.apply_groups_device(d_in.ptr, d_out.ptr, local_bits, base, 0, plan.n_groups, 0)
    got = krylov.download(ld, d_out.ptr)
    assert helpers.close(got, want, 1e-13)
    # accumulate flag: a second application doubles the result
    plan.apply_groups_device(d_in.ptr, d_out.ptr, local_bits, base, 0, plan.n_groups, 1)
    assert helpers.close(krylov.download(ld, d_out.ptr), 2 * want, 1e-13)


def test_diagonal_coulomb_hamiltonian_route():
    # sparse_tools_test.py:1876-1890: get_sparse_operator(DiagonalCoulombHamiltonian)
    data = helpers.golden('fermion_cases.npz')
    dch = hamiltonians.DiagonalCoulomb(data['dch/one_body'], data['dch/two_body'], complex(data['dch/constant']).real)
    mat = get_sparse_operator(dch)
    helpers.assert_csc_matches(mat, data['dch/indptr'], data['dch/indices'], data['dch/data'], RTOL)


# ---- BASELINE.json full sizes: size-independent properties, device-resident ----------------------
def _device_random(n, seed):
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import krylov
    dim = 1 << n
    buf = _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, seed, False, buf.ptr)
    krylov.scal(dim, 1.0 / krylov.nrm2(dim, buf.ptr), buf.ptr)
    return buf


def test_full_size_molecular_28_qubits_properties():
    """Config 4 at full size (T = 88 495, 2^28 amplitudes): Hermiticity, linearity and the
    fused expectation against K1 + dot, all on device-resident vectors."""
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import krylov
    n = 28
    dim = 1 << n
    op = hamiltonians.molecular_jw(14, seed=1234)
    plan = LinearQubitOperator(op, n).plan
    assert (plan.n_terms, plan.n_groups) == (88495, 10466)
    u, v = _device_random(n, 1), _device_random(n, 2)
    hu, hv = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    plan.apply_device(u.ptr, hu.ptr)
    plan.apply_device(v.ptr, hv.ptr)
    scale = float(numpy.sum(numpy.abs(op.coeff)))
    uhv, vhu = krylov.dotc(dim, u.ptr, hv.ptr), krylov.dotc(dim, v.ptr, hu.ptr)
    assert abs(uhv - numpy.conj(vhu)) <= 1e-12 * scale
    e_u = krylov.dotc(dim, u.ptr, hu.ptr)
    assert abs(e_u.imag) <= 1e-12 * scale
    assert abs(plan.expectation_device(u.ptr) - e_u) <= 1e-12 * scale
    # linearity: H(0.3 u - 1.7i v) = 0.3 Hu - 1.7i Hv   (re-uses u and hu as work space)
    a, b = 0.3, -1.7j
    krylov.scal(dim, a, u.ptr)
    krylov.axpy(dim, b, v.ptr, u.ptr)
    krylov.scal(dim, a, hu.ptr)
    krylov.axpy(dim, b, hv.ptr, hu.ptr)      # hu <- expected
    plan.apply_device(u.ptr, hv.ptr)          # hv <- H(combination)
    krylov.axpy(dim, -1.0, hu.ptr, hv.ptr)
    assert krylov.maxabs(dim, hv.ptr) <= 1e-12 * scale


def test_maximum_single_gpu_size_32_qubits_columns():
    """2^32 amplitudes on one GPU (the 32-bit local index boundary): H e_c for basis states c
    near both ends of the index range must equal the column of H computed with Python
    integers -- right rows, right values, nothing written anywhere else."""
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import krylov
    n = 32
    dim = 1 << n
    op = hamiltonians.hubbard_jw(4, 4)
    x, z, c = compiler.compile_qubit_operator(op, n)
    plan = LinearQubitOperator(op, n).plan
    d_in, d_out = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    one = numpy.array([1.0 + 0.0j])
    for col in (dim - 1, (1 << 31) + 12345, 0x5A5A5A5A, 7):
        _lib.call('ofb_memset', _lib.c_void_p(d_in.ptr), _lib.c_int(0), _lib.c_size_t(dim * 16), None)
        krylov.upload(one, d_in.ptr + col * 16)
        plan.apply_device(d_in.ptr, d_out.ptr)
        want = {}
        for xt, zt, ct in zip(x.tolist(), z.tolist(), c.tolist()):
            value = ct * (1j) ** bin(xt & zt).count('1') * (-1) ** bin(col & zt).count('1')
            want[col ^ xt] = want.get(col ^ xt, 0) + value
        for row, value in want.items():
            got = krylov.download(1, d_out.ptr + row * 16)[0]
            assert abs(got - value) <= 1e-13 * max(1.0, abs(value)), (col, row)
        norm2 = krylov.dotc(dim, d_out.ptr, d_out.ptr).real
        assert abs(norm2 - sum(abs(v) ** 2 for v in want.values())) <= 1e-10 * max(norm2, 1.0)
