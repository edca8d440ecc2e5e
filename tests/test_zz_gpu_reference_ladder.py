"""CSC parity ladder against the REAL reference at the largest sizes it can build here
(molecular n <= 14, Hubbard n <= 16; tests/golden/csc_ladder_cases.npz,
written by oracle/make_golden.py csc_ladder): nnz, index width, indptr and indices bit-exact by
sha256; values within 1e-12 of max |value| on sampled entries and, beyond the bar, bit-exact by
sha256 as well (signed zeros included).  Runs last in the suite."""
import hashlib

import numpy
import pytest

import helpers
from openfermion_b200.linalg import jordan_wigner_sparse, qubit_operator_sparse

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(numpy.ascontiguousarray(a).tobytes()).hexdigest()


def _cases():
    data = helpers.golden('csc_ladder_cases.npz')
    return [helpers.Case(data, str(n)) for n in data['names']]


@pytest.mark.parametrize('case', _cases(), ids=lambda c: c.name)
def test_csc_matches_reference_hashes(case):
    n = case.n
    if str(case['route']) == 'pauli':
        op = helpers.qubit_case_operator(case)
        mat = qubit_operator_sparse(op, n)
    else:
        op = helpers.fermion_operator_from(case['terms'], case['coeffs'], case['coeff_is_complex'])
        mat = jordan_wigner_sparse(op, n)
    assert mat.shape == (1 << n, 1 << n) and mat.dtype == numpy.complex128
    assert mat.nnz == int(case['nnz'])
    assert str(mat.indices.dtype) == str(case['index_dtype']) == str(mat.indptr.dtype)
    assert numpy.array_equal(numpy.diff(mat.indptr)[:256], case['colcount_head'])
    assert _sha(mat.indptr) == str(case['indptr_sha256'])
    pos = case['sample_position']
    assert numpy.array_equal(mat.indices[pos], case['sample_indices'])
    assert _sha(mat.indices) == str(case['indices_sha256'])
    scale = numpy.max(numpy.abs(case['sample_data']))
    assert numpy.max(numpy.abs(mat.data[pos] - case['sample_data'])) <= 1e-12 * scale
    total = float(case['data_abs_sum'])
    assert abs(numpy.abs(mat.data).sum() - total) <= 1e-12 * total
    # beyond the parity bar (values within 1e-12): the scipy-order mode reproduces the reference's
    # values bit for bit -- summation order (std::sort replay) AND the signs of zero parts (replay
    # of scipy.sparse.kron's multiplication chain, csrc/ofb_internal.cuh)
    assert _sha(mat.data) == str(case['data_sha256']), 'CSC values differ from the reference in some bit'


# ---- eigenvalue ladder -----------------------------------------------------------------------
def _energy_cases():
    data = helpers.golden('energy_ladder_cases.npz')
    return [helpers.Case(data, str(n)) for n in data['names']]


@pytest.mark.parametrize('case', _energy_cases(), ids=lambda c: c.name)
def test_energies_match_reference_runs(case):
    """tests/golden/energy_ladder_cases.npz (oracle/make_golden.py energy_ladder): the real
    reference's get_ground_state (ARPACK, CSC route and matrix-free), get_gap, variance and
    jw_get_ground_state_at_particle_number at 12-16 qubits.  Ground-state energies within
    1e-10 (BASELINE.json), gaps within 1e-8.  Degenerate spectra: energies only."""
    from openfermion_b200.linalg import (LinearQubitOperator, get_gap, get_ground_state,
                                         get_sparse_operator, jw_get_ground_state_at_particle_number,
                                         variance)
    n = case.n
    op = helpers.qubit_case_operator(case)
    e0 = float(case['e0'])
    mat = get_sparse_operator(op, n)
    energy, state = get_ground_state(mat)
    assert abs(energy - e0) <= 1e-10 * max(1.0, abs(e0))
    assert state.shape == (1 << n,) and abs(numpy.linalg.norm(state) - 1.0) < 1e-10
    if n <= 12 or 'hubbard' in case.name:  # H^2 of the 14-qubit molecular matrix is gigabytes
        assert abs(variance(mat, state)) < 1e-7
    energy_free, _ = get_ground_state(LinearQubitOperator(op, n))
    assert abs(energy_free - e0) <= 1e-10 * max(1.0, abs(e0))
    if case.has('e0_matrix_free'):  # the reference's own two routes agree to this level
        assert abs(float(case['e0_matrix_free']) - e0) < 1e-9
    assert abs(get_gap(mat) - float(case['gap'])) < 1e-8
    for number, want in zip(case['particle_numbers'], case['e0_at_particle_number']):
        got, vec = jw_get_ground_state_at_particle_number(mat, int(number))
        assert abs(got - float(want)) <= 1e-9 * max(1.0, abs(float(want)))
        assert vec.shape == (1 << n,)
        got_symbolic, _ = jw_get_ground_state_at_particle_number(op, int(number), expand=False)
        assert abs(got_symbolic - float(want)) <= 1e-9 * max(1.0, abs(float(want)))


def test_qubit_davidson_matches_reference_run_at_12_qubits():
    """The reference's QubitDavidson(...).get_lowest_n(2) on the 12-qubit molecular Hamiltonian
    (its own test size, davidson_test.py) converged to these two values; eps = 1e-6."""
    import copy
    from openfermion_b200.linalg import QubitDavidson
    case = helpers.Case(helpers.golden('energy_ladder_cases.npz'), 'molecular_6_n12')
    assert bool(case['davidson_success'])
    op = copy.deepcopy(helpers.qubit_case_operator(case))
    op.compress()
    numpy.random.seed(7)
    success, values, vectors = QubitDavidson(op, 12).get_lowest_n(2, max_iterations=150)
    assert success
    assert numpy.max(numpy.abs(numpy.sort(values) - numpy.sort(case['davidson_lowest2']))) < 1e-5
    assert abs(min(values) - float(case['e0'])) < 1e-5
    assert vectors.shape == (1 << 12, 2)
