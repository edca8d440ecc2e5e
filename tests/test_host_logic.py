"""CPU: host-side logic (term compiler, stand-in operators, generators, API surface and
error behaviour) and the C-ABI export check.  No compute calls: there is no GPU here."""
import ctypes
import os
import re

import numpy
import pytest
import scipy.sparse

import helpers
import pauli_oracle as po
from openfermion_b200 import _lib, compiler, hamiltonians
from openfermion_b200.ops import FermionOperator, QubitOperator, count_qubits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ---------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'ofb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ofb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 50
    for name in names:
        assert hasattr(lib, name), name


def test_ctypes_table_covers_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_library_loads_and_reports_version():
    lib = _lib.load()
    assert lib.ofb_version() == 100
    assert lib.ofb_launch_count() >= 0


def test_invalid_plan_arguments_are_rejected_without_gpu():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    x = numpy.array([1 << 5], dtype=numpy.uint64)
    z = numpy.zeros(1, dtype=numpy.uint64)
    c = numpy.ones(1, dtype=numpy.complex128)
    status = lib.ofb_plan_create(3, 1, _lib.ptr(x), _lib.ptr(z), _lib.ptr(c), 0, ctypes.byref(handle))
    assert status == 1  # OFB_ERR_INVALID: term acts outside n_qubits
    assert b'outside n_qubits' in lib.ofb_last_error()


@pytest.mark.skipif(_lib.device_count() > 0, reason='only meaningful without a GPU')
def test_product_path_fails_loudly_without_gpu():
    from openfermion_b200.linalg import LinearQubitOperator, qubit_operator_sparse
    op = LinearQubitOperator(QubitOperator('X0'))  # construction needs no device
    assert op.shape == (2, 2) and op.dtype == numpy.dtype(complex)
    with pytest.raises(_lib.OfbError):
        op * numpy.ones(2)
    with pytest.raises(_lib.OfbError):
        qubit_operator_sparse(QubitOperator('X0'))


# ---- stand-in operators -----------------------------------------------------------
def test_qubit_operator_parsing_and_products():
    assert QubitOperator('X0 Y1').terms == {((0, 'X'), (1, 'Y')): 1.0}
    assert QubitOperator(((1, 'Z'), (0, 'X')), 2.0).terms == {((0, 'X'), (1, 'Z')): 2.0}
    assert (QubitOperator('X0') * QubitOperator('Y0')).terms == {((0, 'Z'),): 1j}
    assert (QubitOperator('Y0') * QubitOperator('X0')).terms == {((0, 'Z'),): -1j}
    assert (QubitOperator('Z0') * QubitOperator('X0')).terms == {((0, 'Y'),): 1j}
    assert (QubitOperator('Y3') * QubitOperator('Y3')).terms == {(): 1.0}
    total = QubitOperator('X0', 1.0) + QubitOperator('X0', -1.0)
    assert total.terms == {}
    assert count_qubits(QubitOperator('Z5')) == 6 and count_qubits(QubitOperator()) == 0


def test_operator_groups_match_reference_slicing():
    op = QubitOperator()
    for q in range(7):
        op += QubitOperator('Z%d' % q, q + 1.0)
    groups = list(op.get_operator_groups(3))
    assert [len(g.terms) for g in groups] == [len(t) for t in po.split_terms(op.terms, 3)]
    merged = QubitOperator()
    for g in groups:
        merged += g
    assert merged == op


def test_fermion_operator_parsing():
    assert FermionOperator('3^ 1').terms == {((3, 1), (1, 0)): 1.0}
    assert count_qubits(FermionOperator('3^ 1')) == 4


# ---- compiler -----------------------------------------------------------------------
def test_mask_folding_follows_reference_bit_convention():
    x, z, c = compiler.compile_qubit_terms(QubitOperator('X0 Y1 Z3', 2.0).terms, 4)
    assert (int(x[0]), int(z[0])) == (0b1100, 0b0101)
    for term in [(), ((0, 'Y'),), ((1, 'Z'), (2, 'X'))]:
        xs, zs, _ = compiler.compile_qubit_terms({term: 1.0}, 3)
        assert (int(xs[0]), int(zs[0])) == po.term_masks(term, 3)[:2]


def test_compiler_rejects_symbolic_coefficients_and_small_n():
    sympy = pytest.importorskip('sympy')
    with pytest.raises(TypeError):
        compiler.compile_qubit_terms({((0, 'X'),): sympy.Symbol('t')}, 1)
    with pytest.raises(ValueError):
        compiler.compile_qubit_terms({((3, 'X'),): 1.0}, 2)


def _dense_from_rows(rows, n):
    dim = 1 << n
    mat = numpy.zeros((dim, dim), dtype=complex)
    cols = numpy.arange(dim, dtype=numpy.uint64)
    for om, ov, x, z, c in zip(*rows):
        ok = (cols & om) == ov
        sign = 1.0 - 2.0 * po.bit_parity(cols & z).astype(float)
        mat[(cols ^ x)[ok].astype(int), cols[ok].astype(int)] += c * sign[ok]
    return mat


@pytest.mark.parametrize('case', helpers.fermion_cases(), ids=lambda c: c.name)
def test_projected_rows_reproduce_reference_fermion_matrices(case):
    op = helpers.fermion_operator_from(case['terms'], case['coeffs'])
    rows = compiler.compile_fermion_terms(op.terms, case.n)
    import scipy.sparse
    want = scipy.sparse.csc_matrix((case['data'], case['indices'], case['indptr']),
                                   shape=(2**case.n, 2**case.n)).toarray()
    got = _dense_from_rows(rows, case.n)
    assert numpy.max(numpy.abs(got - want)) <= 1e-12 * numpy.max(numpy.abs(want))
    # structure: every stored reference entry is produced and nothing else is non-zero
    assert numpy.array_equal(numpy.abs(got) > 1e-13, want != 0)


def test_dead_fermion_terms_emit_nothing():
    rows = compiler.compile_fermion_terms({((0, 1), (0, 1)): 1.0, ((1, 0), (1, 0)): 1.0,
                                           ((0, 1), (0, 0)): 0.0}, 2)
    assert len(rows[0]) == 0
    rows = compiler.compile_fermion_terms({((0, 1), (0, 0)): 2.0}, 2)  # number operator
    assert (int(rows[0][0]), int(rows[1][0]), int(rows[2][0])) == (0b10, 0b10, 0)


# ---- generators -----------------------------------------------------------------------
@pytest.mark.parametrize('name,builder', [
    ('hubbard_2x2_jw', lambda: hamiltonians.hubbard_jw(2, 2, 1.0, 4.0)),
    ('hubbard_2x2_mu_h_jw', lambda: hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2)),
    ('hubbard_2x3_jw', lambda: hamiltonians.hubbard_jw(2, 3, 1.0, 4.0)),
    ('molecular_8_jw', lambda: hamiltonians.molecular_jw(4, seed=1234)),
    ('hubbard_2x2_bk', lambda: hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, encoding='bk')),
    ('molecular_6_bk', lambda: hamiltonians.molecular_jw(3, seed=5, encoding='bk')),
])
def test_generators_reproduce_reference_operators(name, builder):
    data = helpers.golden('qubit_cases.npz')
    ref = helpers.qubit_case_operator(helpers.Case(data, name))
    ours = builder().terms
    assert set(ours) == set(ref.terms)
    assert max(abs(ours[k] - ref.terms[k]) for k in ours) < 1e-13


def test_baseline_config_statistics():
    # T and G of the BASELINE.json configs
    for op, terms, groups in ((hamiltonians.hubbard_jw(3, 4), 133, 49), (hamiltonians.hubbard_jw(4, 4), 177, 65),
                              (hamiltonians.molecular_jw(10), 22351, 2536)):
        assert (len(op), op.n_distinct_x()) == (terms, groups)


def test_mask_operator_index_masks_match_compiler():
    op = hamiltonians.molecular_jw(3, seed=5)
    x1, z1 = op.index_masks()
    x2, z2, c2 = compiler.compile_qubit_terms(op.terms, op.n_qubits)
    assert numpy.array_equal(x1, x2) and numpy.array_equal(z1, z2)
    assert numpy.allclose(op.coeff, c2)


def test_spin_orbital_tensors_layout():
    rng = numpy.random.default_rng(0)
    one, two = rng.normal(size=(2, 2)), rng.normal(size=(2, 2, 2, 2))
    o, t = hamiltonians.spin_orbital_tensors(one, two)
    assert o[2, 0] == one[1, 0] and o[3, 1] == one[1, 0] and o[0, 1] == 0
    assert t[0, 3, 1, 2] == two[0, 1, 0, 1] and t[1, 2, 0, 3] == two[0, 1, 0, 1]
    assert t[2, 2, 0, 0] == two[1, 1, 0, 0] and t[0, 1, 0, 1] == 0


# ---- API surface ------------------------------------------------------------------------
def test_linalg_api_surface_and_errors():
    from openfermion_b200 import linalg
    for name in ['get_sparse_operator', 'qubit_operator_sparse', 'jordan_wigner_sparse',
                 'LinearQubitOperator', 'ParallelLinearQubitOperator', 'LinearQubitOperatorOptions',
                 'generate_linear_qubit_operator', 'get_linear_qubit_operator_diagonal', 'expectation',
                 'variance', 'get_ground_state', 'get_gap', 'inner_product', 'Davidson',
                 'DavidsonOptions', 'QubitDavidson', 'SparseDavidson', 'DavidsonError',
                 'generate_random_vectors', 'append_random_vectors', 'orthonormalize',
                 'jw_number_indices', 'jw_sz_indices', 'jw_number_restrict_operator',
                 'jw_sz_restrict_operator', 'jw_number_restrict_state', 'jw_sz_restrict_state',
                 'jw_get_ground_state_at_particle_number', 'get_number_preserving_sparse_operator',
                 'jw_configuration_state', 'jw_hartree_fock_state', 'get_density_matrix',
                 'eigenspectrum', 'sparse_eigenspectrum', 'SectorLinearQubitOperator']:
        assert hasattr(linalg, name), name
    with pytest.raises(ValueError):
        linalg.LinearQubitOperator(QubitOperator('X3'), 1)
    with pytest.raises(ValueError):
        linalg.qubit_operator_sparse(QubitOperator('X3'), 1)
    with pytest.raises(ValueError):
        linalg.get_linear_qubit_operator_diagonal(QubitOperator('X3'), 1)
    with pytest.raises(TypeError):
        linalg.get_sparse_operator(1.0)
    with pytest.raises(ValueError):
        linalg.LinearQubitOperatorOptions(0)
    with pytest.raises(ValueError):
        linalg.DavidsonOptions(max_subspace=2)
    with pytest.raises(ValueError):
        linalg.DavidsonOptions(eps=0)
    with pytest.raises(ValueError):
        linalg.expectation(linalg.LinearQubitOperator(QubitOperator('X0')), [1, 0])
    opts = linalg.DavidsonOptions()
    assert (opts.max_subspace, opts.max_iterations, opts.eps, opts.real_only) == (100, 300, 1e-6, False)
    opts.set_dimension(10)
    assert opts.max_subspace == 11
    par = linalg.ParallelLinearQubitOperator(QubitOperator('X0') + QubitOperator('Z1'), 2,
                                             linalg.LinearQubitOperatorOptions(2))
    assert par.options.pool is None and len(par.qubit_operator_groups) <= 2
    assert len(par.linear_operators) == len(par.qubit_operator_groups)
    assert isinstance(linalg.generate_linear_qubit_operator(QubitOperator('X0')), linalg.LinearQubitOperator)
    assert linalg.inner_product(numpy.array([1j, 0]), numpy.array([1j, 0])) == 1


def test_small_state_helpers_follow_reference():
    # sparse_tools_test.py known answers: jw_hartree_fock_state(2, 4) is |1100> = index 12
    from openfermion_b200 import linalg
    hf = linalg.jw_hartree_fock_state(2, 4)
    assert hf.dtype == numpy.float64 and list(hf.nonzero()[0]) == [12] and hf[12] == 1.0
    conf = linalg.jw_configuration_state([0, 3], 4)
    assert list(conf.nonzero()[0]) == [9]
    rho = linalg.get_density_matrix([numpy.array([1, 0]), numpy.array([0, 1])], [0.25, 0.75])
    assert scipy.sparse.isspmatrix_csc(rho) and numpy.allclose(rho.toarray(), numpy.diag([0.25, 0.75]))
    spectrum = linalg.sparse_eigenspectrum(scipy.sparse.csc_matrix(numpy.array([[2.0, 1j], [-1j, 2.0]])))
    assert numpy.allclose(spectrum, [1.0, 3.0])
    nonherm = linalg.sparse_eigenspectrum(scipy.sparse.csc_matrix(numpy.array([[0.0, 1.0], [0.0, 0.0]])))
    assert numpy.allclose(nonherm, [0.0, 0.0])
    restricted = linalg.jw_number_restrict_state(numpy.arange(16), 2)
    assert list(restricted) == [3, 5, 9, 6, 10, 12]
    assert list(linalg.jw_sz_restrict_state(numpy.arange(16), 0.0, 2)) == linalg.jw_sz_indices(0.0, 4, 2)


# ---- sector helpers -------------------------------------------------------------
def test_jw_sector_indices_definitions():
    from openfermion_b200.linalg import jw_number_indices, jw_sz_indices
    for n_qubits in (4, 6):
        for k in range(n_qubits + 1):
            got = jw_number_indices(k, n_qubits)
            want = [i for i in range(2**n_qubits) if bin(i).count('1') == k]
            assert sorted(got) == want and len(set(got)) == len(got)
    assert jw_number_indices(2, 4) == [3, 5, 9, 6, 10, 12]  # the reference's combination order

    def sz_of(index, n_qubits):
        up = sum((index >> (n_qubits - 1 - 2 * i)) & 1 for i in range(n_qubits // 2))
        down = sum((index >> (n_qubits - 2 - 2 * i)) & 1 for i in range(n_qubits // 2))
        return (up - down) / 2.0, up + down

    for sz in (-1.0, -0.5, 0.0, 0.5, 1.0):
        got = jw_sz_indices(sz, 6)
        want = [i for i in range(64) if sz_of(i, 6)[0] == sz]
        assert sorted(got) == want and len(set(got)) == len(got)
    got = jw_sz_indices(0.5, 6, n_electrons=3)
    assert sorted(got) == [i for i in range(64) if sz_of(i, 6) == (0.5, 3)]
    with pytest.raises(ValueError):
        jw_sz_indices(0.5, 5)
    with pytest.raises(ValueError):
        jw_sz_indices(0.25, 4)
    with pytest.raises(ValueError):
        jw_sz_indices(0.5, 4, n_electrons=2)


@pytest.mark.needs_reference
def test_jw_sector_indices_match_reference_order():
    import ref_shim
    ref_shim.load()
    from openfermion.linalg.sparse_tools import jw_number_indices as ref_n, jw_sz_indices as ref_sz
    from openfermion_b200.linalg import jw_number_indices, jw_sz_indices
    assert jw_number_indices(3, 7) == ref_n(3, 7)
    for sz in (-1.5, -1.0, 0.0, 0.5, 1.0):
        assert jw_sz_indices(sz, 8) == ref_sz(sz, 8)
    assert jw_sz_indices(-1.0, 8, n_electrons=4) == ref_sz(-1.0, 8, n_electrons=4)


# ---- expectation_computational_basis_state (sparse_tools_test.py:674-740) --------------------
def test_expectation_computational_basis_state_known_answers():
    import scipy.sparse
    from openfermion_b200.linalg import expectation_computational_basis_state as ecbs
    from openfermion_b200.ops import FermionOperator, QubitOperator

    def column(index):
        return scipy.sparse.csc_matrix(([1], ([index], [0])), shape=(16, 1))
    single = FermionOperator('3^ 3', 1.9) + FermionOperator('2^ 1')
    two = FermionOperator('2^ 2', 1.9) + FermionOperator('2^ 1') + FermionOperator('2^ 1^ 2 1', -1.7)
    identity = FermionOperator((), 1.1)
    assert abs(ecbs(single, column(15)) - 1.9) < 1e-12
    assert abs(ecbs(two, column(6)) - 3.6) < 1e-12
    assert abs(ecbs(identity, column(6)) - 1.1) < 1e-12
    assert abs(ecbs(single, [1, 1, 1, 1]) - 1.9) < 1e-12
    assert abs(ecbs(two, [0, 1, 1]) - 3.6) < 1e-12
    assert abs(ecbs(identity, [0, 1, 1]) - 1.1) < 1e-12
    with pytest.raises(TypeError):
        ecbs('never', column(6))
    with pytest.raises(NotImplementedError):
        ecbs(QubitOperator(), column(6))
    bad_order = FermionOperator('2^ 2', 1.9) + FermionOperator('2^ 1') + FermionOperator('2^ 1 2 1^', -1.7)
    with pytest.raises(ValueError):
        ecbs(bad_order, [0, 1, 1])


def test_gate_matrix_helpers_match_reference():
    # jw_sparse_givens_rotation / jw_sparse_particle_hole_transformation_last_mode
    # (sparse_tools.py:516-554; sparse_tools_test.py:565-575 for the errors): golden matrices
    # written by the reference, explicit zeros (theta = 0) included
    import helpers
    from openfermion_b200.linalg import (jw_sparse_givens_rotation,
                                         jw_sparse_particle_hole_transformation_last_mode)
    data = helpers.golden('ladder_cases.npz')
    for k, (i, j, theta, phi, n) in enumerate(data['givens_args']):
        mat = jw_sparse_givens_rotation(int(i), int(j), theta, phi, int(n))
        assert mat.format == 'csc' and mat.nnz == 6 * 2 ** (int(n) - 2)
        helpers.assert_csc_matches(mat, data['givens%d/indptr' % k], data['givens%d/indices' % k],
                                   data['givens%d/data' % k], rtol=0, signed_zeros=False)
        assert numpy.array_equal(mat.data, data['givens%d/data' % k])
    for n in (1, 3, 5):
        mat = jw_sparse_particle_hole_transformation_last_mode(n)
        helpers.assert_csc_matches(mat, data['particle_hole%d/indptr' % n], data['particle_hole%d/indices' % n],
                                   data['particle_hole%d/data' % n], rtol=0, signed_zeros=False)
    with pytest.raises(ValueError):
        jw_sparse_givens_rotation(0, 2, 1.0, 1.0, 5)
    with pytest.raises(ValueError):
        jw_sparse_givens_rotation(4, 5, 1.0, 1.0, 5)
