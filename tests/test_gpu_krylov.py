"""GPU: device-resident Lanczos (get_ground_state / get_gap) and Davidson, against the
reference's own known answers (sparse_tools_test.py:593-604, davidson_test.py) and dense
diagonalisation of oracle matrices.  Energy bar: 1e-10 Ha."""
import warnings

import numpy
import pytest
import scipy.sparse
import scipy.sparse.linalg

import helpers
import pauli_oracle as po
from openfermion_b200 import hamiltonians
from openfermion_b200.linalg import (Davidson, DavidsonOptions, LinearQubitOperator, QubitDavidson,
                                     SparseDavidson, append_random_vectors, generate_random_vectors,
                                     get_gap, get_ground_state, get_sparse_operator, orthonormalize,
                                     qubit_operator_sparse, expectation)
from openfermion_b200.ops import QubitOperator

pytestmark = pytest.mark.gpu


def _dense(op, n):
    return po.qubit_operator_sparse(op.terms, n).toarray()


# ---- get_ground_state / get_gap -----------------------------------------------------------
def test_ground_state_known_answer():
    op = QubitOperator('Y0 X1') + QubitOperator('Z0 Z1')
    energy, state = get_ground_state(get_sparse_operator(op))
    expected = numpy.array([0, 1j, 1, 0]) / numpy.sqrt(2.0)
    assert abs(energy + 2) < 1e-10
    assert abs(abs(numpy.vdot(expected, state)) - 1.0) < 1e-8
    energy2, state2 = get_ground_state(LinearQubitOperator(op))
    assert abs(energy2 + 2) < 1e-10 and abs(abs(numpy.vdot(expected, state2)) - 1.0) < 1e-8


@pytest.mark.parametrize('builder', [lambda: hamiltonians.hubbard_jw(2, 2), lambda: hamiltonians.molecular_jw(4, 1234),
                                     lambda: hamiltonians.hubbard_jw(2, 3, 1.0, 4.0, 0.3, 0.2)])
def test_ground_state_matches_dense_diagonalisation(builder):
    op = builder()
    n = op.n_qubits
    want = numpy.linalg.eigvalsh(_dense(op, n))
    lin = LinearQubitOperator(op, n)
    energy, state = get_ground_state(lin)
    assert isinstance(energy, float) and state.shape == (2**n,) and state.dtype == numpy.complex128
    assert abs(energy - want[0]) < 1e-10
    resid = lin * state - energy * state
    assert numpy.linalg.norm(resid) < 1e-7
    assert abs(numpy.linalg.norm(state) - 1.0) < 1e-12
    assert abs(expectation(lin, state) - energy) < 1e-10
    energy_csc, _ = get_ground_state(qubit_operator_sparse(op, n))
    assert abs(energy_csc - want[0]) < 1e-10


def test_ground_state_with_initial_guess():
    op = hamiltonians.hubbard_jw(2, 2)
    want = numpy.linalg.eigvalsh(_dense(op, 8))[0]
    guess = numpy.random.rand(256) + 0j
    energy, _ = get_ground_state(LinearQubitOperator(op, 8), guess)
    assert abs(energy - want) < 1e-10


def test_lih_ground_state_energy():
    # config 1: testing/lih_integration_test.py:70-71,103-106
    data = helpers.golden('lih_sto3g.npz')
    ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                       data['two_body_integrals'])
    csc = get_sparse_operator(ham)
    energy, state = get_ground_state(csc)
    assert abs(energy - float(data['reference_e0'])) < 1e-10
    assert abs(energy - float(data['fci_energy'])) < 1e-7
    assert abs(expectation(csc, state) - energy) < 1e-10
    qop = helpers.qubit_operator_from(data['qubit_terms'], data['qubit_coeffs'])
    energy_mf, _ = get_ground_state(LinearQubitOperator(qop, 12))
    assert abs(energy_mf - float(data['reference_e0'])) < 1e-10


def test_get_gap():
    # sparse_tools_test.py:1114-1124
    op = QubitOperator('Y0 X1') + QubitOperator('Z0 Z1')
    assert abs(get_gap(get_sparse_operator(op)) - 2.0) < 1e-9
    with pytest.raises(ValueError):
        get_gap(get_sparse_operator(QubitOperator('X0 Y1') + QubitOperator('Z0 Z1', 1j)))
    hub = hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2)
    vals = numpy.linalg.eigvalsh(_dense(hub, 8))
    assert abs(get_gap(LinearQubitOperator(hub, 8)) - (vals[1] - vals[0])) < 1e-8


# ---- Davidson: the reference's trajectory-pinning tests (davidson_test.py:120-311) --------
def _reference_matrix(dimension=10):
    numpy.random.seed(dimension)
    rand = numpy.array(numpy.random.rand(dimension, dimension))
    numpy.random.seed(dimension * 2)
    diag = numpy.array(range(dimension)) + numpy.random.rand(dimension)
    return rand + rand.conj().T + numpy.diag(diag)


EIGS10 = numpy.array([1.15675714, 1.59132505, 2.62268014, 4.44533793, 5.3722743, 5.54393114,
                      7.73652405, 8.50089897, 9.4229309, 15.54405993])


def _davidson10():
    matrix = _reference_matrix()
    lin = scipy.sparse.linalg.LinearOperator((10, 10), matvec=lambda v: numpy.dot(matrix, v))
    return Davidson(lin, numpy.diag(matrix)), matrix, numpy.eye(10, 10)


def test_davidson_init_and_argument_checks():
    dav, matrix, guess = _davidson10()
    assert dav.options.max_subspace == 11
    with pytest.raises(ValueError):
        Davidson(matrix, numpy.diag(matrix))
    with pytest.raises(ValueError):
        dav.get_lowest_n(0)
    with pytest.raises(ValueError):
        dav.get_lowest_n(1, numpy.ones((20, 1), dtype=complex))
    with pytest.raises(ValueError):
        dav.get_lowest_n(1, numpy.zeros((10, 1), dtype=complex))


def test_davidson_two_iterations_not_converged():
    dav, _, guess = _davidson10()
    success, values, _ = dav.get_lowest_n(1, guess[:, :1], max_iterations=2)
    assert not success and numpy.allclose(values, [1.41556103])


def test_davidson_default_and_one():
    dav, _, guess = _davidson10()
    numpy.random.seed(10)
    success, values, _ = dav.get_lowest_n()
    assert success and numpy.allclose(values, EIGS10[:1])
    success, values, _ = dav.get_lowest_n(1, guess[:, :1], max_iterations=10)
    assert success and numpy.allclose(values, EIGS10[:1])


def test_davidson_two_converges_in_five():
    dav, matrix, guess = _davidson10()
    success, values, vectors = dav.get_lowest_n(2, guess[:, :2], max_iterations=5)
    assert success and numpy.allclose(values, EIGS10[:2])
    assert numpy.allclose(matrix @ vectors, vectors * values)


def test_davidson_two_small_subspace_restart_path():
    dav, matrix, guess = _davidson10()
    dav.options.max_subspace = 8
    success, values, vectors = dav.get_lowest_n(2, guess[:, :2], max_iterations=5)
    assert not success and numpy.allclose(values, [1.1572995, 1.61393264])
    assert not numpy.allclose(matrix @ vectors, vectors * values)


def test_davidson_six_and_all():
    dav, _, guess = _davidson10()
    success, values, _ = dav.get_lowest_n(6, guess[:, :6], max_iterations=2)
    assert success and numpy.allclose(values, EIGS10[:6])
    success, values, _ = dav.get_lowest_n(10, guess[:, :10], max_iterations=1)
    assert success and numpy.allclose(values, EIGS10)


# ---- QubitDavidson (davidson_test.py:314-516), 12 qubits ---------------------------------------
def _difference(lin, values, vectors):
    return numpy.max(numpy.abs(lin.matmat(vectors) - vectors * values))


def test_qubit_davidson_z_sum():
    n, dim = 12, 2**12
    op = QubitOperator()
    for i in range(4):
        numpy.random.seed(dim + i)
        op += QubitOperator(((i, 'Z'),), numpy.random.rand(1)[0])
    op *= 2
    dav = QubitDavidson(op, n)
    numpy.random.seed(dim)
    success, values, vectors = dav.get_lowest_n(6, numpy.random.rand(dim, 6), max_iterations=20)
    assert success and numpy.allclose(values, -3.80376934 * numpy.ones(6))
    assert _difference(dav.linear_operator, values, vectors) < 1e-5


def test_qubit_davidson_zzx_fewer_guesses_than_wanted():
    n, dim = 12, 2**12
    dav = QubitDavidson(QubitOperator('Z0 Z1 X2') * 2, n)
    numpy.random.seed(dim)
    success, values, vectors = dav.get_lowest_n(6, numpy.random.rand(dim, 3), max_iterations=10)
    assert success and numpy.allclose(values, -2 * numpy.ones(6))
    assert _difference(dav.linear_operator, values, vectors) < 1e-5


def test_qubit_davidson_xyz_complex_guess():
    n, dim = 12, 2**12
    dav = QubitDavidson(QubitOperator('X0 Y1 Z3') * 2, n)
    numpy.random.seed(dim)
    guess = 1.0j * numpy.random.rand(dim, 6)
    numpy.random.seed(dim * 2)
    guess += numpy.random.rand(dim, 6)
    success, values, vectors = dav.get_lowest_n(6, guess, max_iterations=10)
    assert success and numpy.allclose(values, -2 * numpy.ones(6))
    assert _difference(dav.linear_operator, values, vectors) < 1e-5


def test_qubit_davidson_real_only():
    n, dim = 12, 2**12
    dav = QubitDavidson(QubitOperator('Z3') * 2, n)
    dav.options.real_only = True
    numpy.random.seed(dim)
    guess = 1.0j * numpy.random.rand(dim, 6)
    numpy.random.seed(dim * 2)
    guess += numpy.random.rand(dim, 6)
    with pytest.warns(RuntimeWarning):
        success, values, vectors = dav.get_lowest_n(6, guess, max_iterations=10)
    assert success and numpy.allclose(values, -2 * numpy.ones(6))
    assert numpy.allclose(numpy.real(vectors), vectors)


def test_qubit_davidson_hubbard_against_dense():
    op = hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2)
    want = numpy.linalg.eigvalsh(_dense(op, 8))[:3]
    numpy.random.seed(3)
    dav = QubitDavidson(op, 8, DavidsonOptions(max_subspace=40, eps=1e-8))
    success, values, vectors = dav.get_lowest_n(3, max_iterations=200)
    assert success and numpy.allclose(values, want, atol=1e-9)


# ---- SparseDavidson (davidson_test.py:519-621) ---------------------------------------------
@pytest.mark.parametrize('complex_matrix', [False, True])
def test_sparse_davidson_against_eigh(complex_matrix):
    dimension = 60
    numpy.random.seed(dimension)
    diagonal = numpy.sort(numpy.random.rand(dimension)) * 30
    off = numpy.random.rand(dimension - 1) * (1j if complex_matrix else 1.0)
    matrix = numpy.diag(diagonal).astype(complex if complex_matrix else float)
    matrix += numpy.diag(off, 1) + numpy.diag(off.conj(), -1)
    want = numpy.linalg.eigvalsh(matrix)[:2]
    for fmt in (scipy.sparse.csc_matrix, scipy.sparse.coo_matrix):
        dav = SparseDavidson(fmt(matrix), DavidsonOptions(max_subspace=30, eps=1e-8))
        numpy.random.seed(1)
        success, values, vectors = dav.get_lowest_n(2, max_iterations=200)
        assert success and numpy.allclose(values, want, atol=1e-8)
        assert numpy.max(numpy.abs(matrix @ vectors - vectors * values)) < 1e-6


# ---- helpers (davidson_test.py:624-707) ----------------------------------------------------
def test_orthonormalize_in_place_and_drops_dependent_columns():
    numpy.random.seed(0)
    vectors = numpy.random.rand(64, 4) + 1j * numpy.random.rand(64, 4)
    vectors[:, 0] /= numpy.linalg.norm(vectors[:, 0])
    vectors[:, 2] = 2 * vectors[:, 0]  # dependent on the first
    original = vectors.copy()
    out = orthonormalize(vectors, 1)
    assert out.shape == (64, 3)
    assert numpy.allclose(out.conj().T @ out, numpy.eye(3), atol=1e-12)
    assert numpy.allclose(out[:, 0], original[:, 0])
    assert numpy.shares_memory(out, vectors)
    want = po.orthonormalize(original.copy(), 1)
    assert numpy.allclose(numpy.abs(want.conj().T @ out), numpy.eye(3), atol=1e-10)


def test_generate_and_append_random_vectors():
    vecs = generate_random_vectors(32, 3)
    assert vecs.shape == (32, 3) and numpy.iscomplexobj(vecs)
    assert numpy.allclose(vecs.conj().T @ vecs, numpy.eye(3), atol=1e-12)
    real = generate_random_vectors(32, 2, real_only=True)
    assert numpy.allclose(real.imag, 0)
    more = append_random_vectors(vecs, 2)
    assert more.shape == (32, 5) and numpy.allclose(more[:, :3], vecs)
    assert numpy.allclose(more.conj().T @ more, numpy.eye(5), atol=1e-12)
    assert append_random_vectors(vecs, 0) is vecs
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        full = append_random_vectors(numpy.eye(3, dtype=complex), 2)
    assert full.shape[1] == 3  # more columns than rows cannot be added


# ---- sector-restricted ground states (sparse_tools_test.py: jw_get_ground_state_at_particle_number)
def test_ground_state_at_particle_number_matches_dense_sector():
    from openfermion_b200.linalg import (jw_get_ground_state_at_particle_number,
                                         jw_number_restrict_operator, jw_sz_restrict_operator)
    op = hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2)
    csc = qubit_operator_sparse(op, 8)
    dense = _dense(op, 8)
    for particles in (0, 1, 2, 4, 8):
        idx = [i for i in range(256) if bin(i).count('1') == particles]
        want = numpy.linalg.eigvalsh(dense[numpy.ix_(idx, idx)])[0]
        energy, state = jw_get_ground_state_at_particle_number(csc, particles)
        assert abs(energy - want) < 1e-10
        assert state.shape == (256,) and abs(numpy.linalg.norm(state) - 1) < 1e-10
        assert abs(numpy.vdot(state, dense @ state) - want) < 1e-9
    restricted = jw_number_restrict_operator(csc, 4)
    assert restricted.shape == (70, 70)
    sz0 = jw_sz_restrict_operator(csc, 0.0, n_electrons=4)
    assert sz0.shape == (36, 36)
    lowest_sz0 = numpy.linalg.eigvalsh(sz0.toarray())[0]
    assert lowest_sz0 >= numpy.linalg.eigvalsh(restricted.toarray())[0] - 1e-10


def test_reductions_from_several_host_threads_and_streams():
    # the vector reductions share one scratch buffer per device; concurrent callers serialise on
    # a per-device lock instead of racing on the partial sums
    import concurrent.futures
    import ctypes
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import krylov
    n = 1 << 18
    rng = numpy.random.default_rng(21)
    vectors = [rng.normal(size=n) + 1j * rng.normal(size=n) for _ in range(4)]
    bufs = [_lib.to_device(v) for v in vectors]
    want = [(numpy.vdot(v, v), numpy.linalg.norm(v), numpy.max(numpy.abs(v))) for v in vectors]

    def work(k):
        _lib.call('ofb_set_device', ctypes.c_int(0))
        stream = ctypes.c_void_p()
        _lib.call('ofb_stream_create', ctypes.byref(stream))
        try:
            for _ in range(50):
                d = krylov.dotc(n, bufs[k].ptr, bufs[k].ptr, stream.value)
                nr = krylov.nrm2(n, bufs[k].ptr, stream.value)
                mx = krylov.maxabs(n, bufs[k].ptr, stream.value)
                assert abs(d - want[k][0]) <= 1e-12 * abs(want[k][0])
                assert abs(nr - want[k][1]) <= 1e-12 * want[k][1]
                assert abs(mx - want[k][2]) <= 1e-14 * want[k][2]
        finally:
            _lib.call('ofb_stream_destroy', stream)
        return True

    with concurrent.futures.ThreadPoolExecutor(4) as pool:
        assert all(pool.map(work, range(4)))
    for b in bufs:
        b.free()
