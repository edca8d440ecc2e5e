"""GPU: parity of the symmetry-sector kernels (csrc/sector.cu, through the Python mirror and
the C ABI) against the golden vectors of the real reference (tests/golden/sector_cases.npz),
the sector oracle, and -- at sizes the reference cannot reach -- size-independent
properties (Hermiticity, agreement with the full-space kernel, known Hubbard energies).
Bars: indices bit-exact, values / matvecs / expectations within 1e-12 relative, energies
within 1e-10."""
import numpy
import pytest
import scipy.sparse

import helpers
import pauli_oracle
import sector_oracle
from helpers import fermion_operator_from, golden, qubit_operator_from
from openfermion_b200 import hamiltonians, sectors
from openfermion_b200.linalg import (LinearQubitOperator, SectorLinearQubitOperator, expectation,
                                     get_ground_state, get_number_preserving_sparse_operator,
                                     jw_get_ground_state_at_particle_number,
                                     jw_number_restrict_operator, jw_sz_restrict_operator)

pytestmark = pytest.mark.gpu
RTOL = 1e-12
DATA = golden('sector_cases.npz')
LIST_NAMES = [str(n) for n in DATA['list_names']]
OPERATORS = [str(n) for n in DATA['names']]
NP_CASES = [str(n) for n in DATA['np_names']]


def _parse_tag(tag):
    kind, ne, sz = str(tag).split('_')
    return kind, (None if ne == 'None' else int(ne)), (None if sz == 'None' else float(sz))


def _operators(name):
    n = int(DATA[name + '/n_qubits'])
    fop = fermion_operator_from(DATA[name + '/fermion_terms'], DATA[name + '/fermion_coeffs'],
                                DATA[name + '/fermion_coeff_is_complex'])
    qop = qubit_operator_from(DATA[name + '/qubit_terms'], DATA[name + '/qubit_coeffs'],
                              DATA[name + '/qubit_coeff_is_complex'])
    return n, fop, qop


def _cases():
    return [(name, str(tag)) for name in OPERATORS for tag in DATA[name + '/sectors']]


@pytest.mark.parametrize('name', LIST_NAMES)
def test_device_basis_enumeration_equals_reference_lists(name):
    parts = name.split('_')
    want = DATA['lists/' + name].astype(numpy.uint64)
    if parts[0] == 'number':
        sector = sectors.Sector.number(int(parts[1]), int(parts[2]))
    else:
        ne = None if parts[3] == 'None' else int(parts[3])
        sector = sectors.Sector.sz(int(parts[1]), float(parts[2]), ne)
    assert sector.dim == len(want)
    assert numpy.array_equal(sector.basis(), want)


def test_custom_spin_map_sector_uses_search_mode():
    want = DATA['lists/sz_blocked_6_0.0_2'].astype(numpy.uint64)
    sector = sectors.Sector.sz(6, 0.0, 2, up_index=lambda i: i, down_index=lambda i: i + 3)
    assert sector.mode == 1 and numpy.array_equal(sector.basis(), want)
    hub = hamiltonians.random_pauli_sum(6, 30, 3, True)
    full = pauli_oracle.qubit_operator_sparse(hub.terms, 6)
    idx = want.astype(numpy.int64)
    rng = numpy.random.default_rng(1)
    v = rng.normal(size=len(idx)) + 1j * rng.normal(size=len(idx))
    got = sector.apply_host(LinearQubitOperator(hub, 6).plan, v)
    assert helpers.close(got, full[numpy.ix_(idx, idx)] @ v, RTOL)


@pytest.mark.parametrize('name,tag', _cases())
def test_sector_matvec_expectation_diagonal(name, tag):
    n, _, qop = _operators(name)
    kind, ne, sz = _parse_tag(tag)
    op = SectorLinearQubitOperator(qop, n, n_electrons=ne, sz_value=sz)
    v = DATA['{}/{}/v'.format(name, tag)]
    want = DATA['{}/{}/Hv'.format(name, tag)]
    assert op.shape == (len(v), len(v))
    got = op * v
    assert got.dtype == numpy.complex128
    assert helpers.close(got, want, RTOL)
    # matmat, two columns
    both = op.matmat(numpy.stack([v, 1j * v], axis=1))
    assert helpers.close(both[:, 0], want, RTOL) and helpers.close(both[:, 1], 1j * want, RTOL)
    e_want = complex(DATA['{}/{}/expectation'.format(name, tag)])
    assert abs(expectation(op, v) - e_want) <= RTOL * max(abs(e_want), 1.0)
    diag = DATA['{}/{}/diagonal'.format(name, tag)]
    assert helpers.close(op.diagonal(), diag.real, RTOL)


@pytest.mark.parametrize('name,tag', _cases())
def test_restricted_csc_from_symbolic_operators(name, tag):
    n, fop, qop = _operators(name)
    kind, ne, sz = _parse_tag(tag)
    if kind == 'number':
        got_f = jw_number_restrict_operator(fop, ne, n)
        got_q = jw_number_restrict_operator(qop, ne, n)
    else:
        got_f = jw_sz_restrict_operator(fop, sz, ne, n)
        got_q = jw_sz_restrict_operator(LinearQubitOperator(qop, n), sz, ne)
    key = '{}/{}/'.format(name, tag)
    # fermion route: structure bit-exact, values to 1e-12
    helpers.assert_csc_matches(got_f, DATA[key + 'f_indptr'], DATA[key + 'f_indices'], DATA[key + 'f_data'])
    # Pauli route: the reference's own full matrix carries cancellation residues
    coeff_abs_sum = float(numpy.sum(numpy.abs(DATA[name + '/qubit_coeffs'])))
    helpers.assert_csc_matches_modulo_residues(got_q, DATA[key + 'q_indptr'], DATA[key + 'q_indices'],
                                               DATA[key + 'q_data'], coeff_abs_sum)


def test_restrict_functions_still_slice_materialised_matrices():
    n, fop, _ = _operators('hubbard_2x2')
    full = pauli_oracle.jordan_wigner_sparse(fop.terms, n)
    got = scipy.sparse.csc_matrix(jw_number_restrict_operator(full, 4))
    key = 'hubbard_2x2/number_4_None/'
    got.sort_indices()
    assert numpy.array_equal(got.indices, DATA[key + 'f_indices'])
    assert numpy.array_equal(got.data, DATA[key + 'f_data'])
    dense = jw_sz_restrict_operator(full.toarray(), 0.0, 4)
    assert dense.shape == (36, 36)


@pytest.mark.parametrize('key', NP_CASES)
def test_get_number_preserving_sparse_operator(key):
    refdet = DATA[key + '/reference_determinant']
    level = int(DATA[key + '/excitation_level'])
    name = str(DATA[key + '/operator'])
    n, fop, _ = _operators(name)
    got = get_number_preserving_sparse_operator(
        fop, n, int(DATA[key + '/num_electrons']), spin_preserving=bool(DATA[key + '/spin_preserving']),
        reference_determinant=None if refdet.size == 0 else list(refdet),
        excitation_level=None if level < 0 else level)
    assert scipy.sparse.isspmatrix_csc(got)
    assert str(got.dtype) == str(DATA[key + '/dtype'])
    helpers.assert_csc_matches(got, DATA[key + '/indptr'], DATA[key + '/indices'], DATA[key + '/data'])


def test_number_preserving_operator_rejects_non_conserving_terms():
    from openfermion_b200.ops import FermionOperator
    with pytest.raises(ValueError, match="doesn't preserve particle number"):
        get_number_preserving_sparse_operator(FermionOperator('2^ 1^ 0'), 4, 2)


@pytest.mark.parametrize('name,ne', [('hubbard_2x2', 2), ('hubbard_2x2', 4), ('molecular_6', 3),
                                     ('molecular_6', 0)])
def test_ground_state_at_particle_number(name, ne):
    n, fop, qop = _operators(name)
    want = float(DATA['{}/gs_number_{}/energy'.format(name, ne)])
    support = set(sector_oracle.jw_number_indices(ne, n))
    # symbolic operator: matrix-free inside the sector
    e, state = jw_get_ground_state_at_particle_number(qop, ne)
    assert abs(e - want) < 1e-10
    assert state.shape == (2**n,) and set(numpy.nonzero(numpy.abs(state) > 1e-9)[0]) <= support
    # materialised matrix: the reference's own call
    full = pauli_oracle.jordan_wigner_sparse(fop.terms, n)
    e2, state2 = jw_get_ground_state_at_particle_number(full, ne)
    assert abs(e2 - want) < 1e-10
    assert abs(numpy.vdot(state2, full @ state2) - want) < 1e-9


@pytest.mark.parametrize('name,tag', [c for c in _cases() if c[0] != 'molecular_4_complex'])
def test_sector_ground_energy_by_device_lanczos(name, tag):
    key = '{}/{}/e0'.format(name, tag)
    if key not in DATA.files:
        pytest.skip('sector too small')
    n, _, qop = _operators(name)
    kind, ne, sz = _parse_tag(tag)
    op = SectorLinearQubitOperator(qop, n, n_electrons=ne, sz_value=sz)
    e, vec = get_ground_state(op)
    assert abs(e - float(DATA[key])) < 1e-10
    assert abs(numpy.vdot(vec, op * vec) - e) < 1e-9


def test_sector_kernel_agrees_with_full_space_kernel_20_qubits():
    # Hubbard 2x5 (20 qubits), 8 electrons, S_z = 0: embed a sector vector, apply the
    # full-space K1 kernel, restrict again
    op = hamiltonians.hubbard_jw(2, 5, 1.0, 4.0)
    lin = LinearQubitOperator(op, 20)
    sec = SectorLinearQubitOperator(lin, n_electrons=8, sz_value=0.0)
    assert sec.shape[0] == 210 * 210
    rng = numpy.random.default_rng(3)
    v = rng.normal(size=sec.shape[0]) + 1j * rng.normal(size=sec.shape[0])
    full = sec.sector.expand_state_host(v)
    want = sec.sector.restrict_state_host(lin * full)
    assert helpers.close(sec * v, want, RTOL)
    # the Hamiltonian conserves N and S_z: nothing leaks out of the sector
    assert abs(numpy.linalg.norm(lin * full) - numpy.linalg.norm(want)) < 1e-9


def test_hubbard_4x4_sector_is_hermitian_and_hits_known_energy():
    # 32 qubits, 10 electrons, S_z = 0: 4368^2 = 1.9e7 amplitudes instead of 2^32.  The
    # closed-shell 10-electron ground energy of the periodic 4x4 lattice at U = 4 is
    # -19.58094 (the full-space 8-GPU Lanczos of config 5 returns the same number).
    op = hamiltonians.hubbard_jw(4, 4, 1.0, 4.0)
    sec = SectorLinearQubitOperator(op, 32, n_electrons=10, sz_value=0.0)
    dim = sec.shape[0]
    assert dim == 4368 ** 2
    rng = numpy.random.default_rng(7)
    u = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    hu, hv = sec * u, sec * v
    a, b = numpy.vdot(u, hv), numpy.vdot(hu, v)
    assert abs(a - b) <= 1e-12 * abs(a)
    e, vec = get_ground_state(sec)
    assert abs(e - (-19.58094)) < 2e-5
    assert numpy.linalg.norm(sec * vec - e * vec) < 1e-6


def test_davidson_inside_a_sector():
    # the generic Davidson class (davidson.py:71) on the sector operator + sector diagonal
    from openfermion_b200.linalg import Davidson, DavidsonOptions
    n, _, qop = _operators('molecular_8')
    op = SectorLinearQubitOperator(qop, n, n_electrons=4, sz_value=0.0)
    want = float(DATA['molecular_8/sz_4_0.0/e0'])
    dav = Davidson(op, op.diagonal(), DavidsonOptions(max_subspace=30, eps=1e-8))
    success, values, vectors = dav.get_lowest_n(1)
    assert success and abs(values[0] - want) < 1e-8
    assert vectors.shape == (op.shape[0], 1)
    assert numpy.linalg.norm(op * vectors[:, 0] - values[0] * vectors[:, 0]) < 1e-6


def test_hubbard_4x4_half_filling_sector_ground_state():
    # 32 qubits, 16 electrons, S_z = 0: 12870^2 = 1.66e8 amplitudes (2.65 GB) on one GPU, where
    # the full space needs 64 GiB per vector.  E0 of the periodic 4x4 lattice at U = 4, half
    # filling: -13.62185 (exact-diagonalisation literature value, -0.85137 per site).
    op = hamiltonians.hubbard_jw(4, 4, 1.0, 4.0)
    sec = SectorLinearQubitOperator(op, 32, n_electrons=16, sz_value=0.0)
    assert sec.shape[0] == 12870 ** 2
    e, vec = jw_get_ground_state_at_particle_number(op, 16, sz_value=0.0, expand=False)
    assert abs(e - (-13.62185)) < 2e-5
    assert vec.shape == (12870 ** 2,) and abs(numpy.linalg.norm(vec) - 1.0) < 1e-9
    # <psi|H|psi> through the fused sector expectation equals the eigenvalue
    assert abs(expectation(sec, vec) - e) < 1e-8


def test_generic_sector_kernel_matches_fast_path(monkeypatch):
    # number / S_z+number sectors normally run k_sector_apply_fast; the generic kernel (n > 32,
    # class tables, search mode) must give the same vector on the same input
    n, _, qop = _operators('molecular_8')
    op = SectorLinearQubitOperator(qop, n, n_electrons=4, sz_value=0.0)
    v = DATA['molecular_8/sz_4_0.0/v']
    fast = op * v
    monkeypatch.setenv('OFB_SECTOR_GENERIC', '1')
    generic = op * v
    assert helpers.close(generic, fast, 1e-14)
    assert helpers.close(generic, DATA['molecular_8/sz_4_0.0/Hv'], RTOL)


def _restricted_dense(terms, n_qubits, basis):
    """Independent host evaluation of P H P on a short basis list: every Pauli string applied to
    every basis ket with integer arithmetic (qubit q = index bit n-1-q)."""
    position = {int(b): k for k, b in enumerate(basis)}
    out = numpy.zeros((len(basis), len(basis)), dtype=complex)
    for term, coefficient in terms.items():
        x = z = 0
        n_y = 0
        for qubit, action in term:
            bit = 1 << (n_qubits - 1 - qubit)
            if action in 'XY':
                x |= bit
            if action in 'YZ':
                z |= bit
            n_y += action == 'Y'
        for col, ket in enumerate(basis):
            ket = int(ket)
            row = position.get(ket ^ x)
            if row is not None:
                # P|ket> = i^{n_y} (-1)^{popc(ket & z)} |ket ^ x>
                out[row, col] += coefficient * (1j ** n_y) * (-1) ** bin(ket & z).count('1')
    return out


@pytest.mark.parametrize('sites', [17, 20])
def test_sectors_beyond_32_qubits_use_the_64_bit_kernel(sites):
    # 1 x 17 / 1 x 20 Hubbard chains = 34 / 40 qubits, 2 electrons, S_z = 0: sites^2 amplitudes
    # out of 2^34 / 2^40 (the enumeration skips high halves that cannot reach the counts)
    n = 2 * sites
    op = hamiltonians.hubbard_jw(1, sites, 1.0, 4.0)
    sec = SectorLinearQubitOperator(op, n, n_electrons=2, sz_value=0.0)
    dim = sites * sites
    assert sec.shape == (dim, dim)
    basis = sec.indices()
    assert list(basis) == sector_oracle.jw_sz_indices(0.0, n, 2)
    dense = _restricted_dense(op.terms, n, basis)
    assert numpy.allclose(dense, dense.conj().T)
    got = sec.matmat(numpy.eye(dim, dtype=complex))
    assert helpers.close(got, dense, RTOL)
    assert helpers.close(sec.diagonal(), numpy.diag(dense).real, RTOL)
    e, _ = get_ground_state(sec)
    assert abs(e - numpy.linalg.eigvalsh(dense)[0]) < 1e-10
    number = SectorLinearQubitOperator(op, n, n_electrons=1)
    assert number.shape == (n, n) and list(number.indices()) == sector_oracle.jw_number_indices(1, n)
    dense1 = _restricted_dense(op.terms, n, number.indices())
    assert helpers.close(number.matmat(numpy.eye(n, dtype=complex)), dense1, RTOL)
    # the restricted CSC of the same symbolic operator
    csc = jw_sz_restrict_operator(op, 0.0, 2, n)
    assert csc.shape == (dim, dim) and helpers.close(csc.toarray(), dense, RTOL)


def test_eigenspectrum_of_symbolic_operators():
    from openfermion_b200.linalg import eigenspectrum
    n, fop, qop = _operators('hubbard_2x2')
    want = numpy.linalg.eigvalsh(pauli_oracle.jordan_wigner_sparse(fop.terms, n).toarray())
    assert numpy.allclose(eigenspectrum(fop, n), want, atol=1e-10)
    assert numpy.allclose(eigenspectrum(qop, n), want, atol=1e-10)
