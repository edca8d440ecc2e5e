"""CPU tests of the symmetry-sector host logic and of the sector oracle against the golden
vectors generated from the real reference (oracle/make_golden.py sector_cases)."""
import numpy
import pytest
import scipy.sparse

import pauli_oracle
import sector_oracle
from helpers import fermion_operator_from, golden
from openfermion_b200 import ops, sectors
from openfermion_b200.linalg import sparse_tools as st

DATA = golden('sector_cases.npz')
LIST_NAMES = [str(n) for n in DATA['list_names']]
OPERATORS = [str(n) for n in DATA['names']]
NP_CASES = [str(n) for n in DATA['np_names']]


def _parse_list(name):
    parts = name.split('_')
    if parts[0] == 'number':
        return 'number', int(parts[1]), int(parts[2]), None
    return 'sz', int(parts[1]), None if parts[3] == 'None' else int(parts[3]), float(parts[2])


def _tables_for(kind, n, ne, sz, lo_bits=None):
    if kind == 'number':
        return sectors.number_sector_tables(n, ne, lo_bits)
    if ne is None:
        return sectors.sz_sector_tables(n, sz, lo_bits)
    return sectors.sz_number_sector_tables(n, sz, ne, lo_bits)


@pytest.mark.parametrize('name', LIST_NAMES)
def test_index_lists_of_oracle_and_mirror_match_reference(name):
    kind, n, ne, sz = _parse_list(name)
    want = DATA['lists/' + name].tolist()
    if kind == 'number':
        assert sector_oracle.jw_number_indices(ne, n) == want
        assert st.jw_number_indices(ne, n) == want
    else:
        assert sector_oracle.jw_sz_indices(sz, n, ne) == want
        assert st.jw_sz_indices(sz, n, ne) == want


@pytest.mark.parametrize('name', LIST_NAMES)
def test_two_table_index_map_reproduces_reference_order(name):
    kind, n, ne, sz = _parse_list(name)
    want = DATA['lists/' + name].astype(numpy.uint64)
    for lo_bits in sorted({n // 2, 0, n, max(0, n // 2 - 1), min(n, n // 2 + 1)}):
        tables = _tables_for(kind, n, ne, sz, lo_bits)
        assert tables.dim == len(want)
        assert numpy.array_equal(tables.basis(), want), (name, lo_bits)
        # membership: exactly the listed states
        everything = numpy.arange(1 << n, dtype=numpy.uint64)
        assert set(everything[tables.member(everything)].tolist()) == set(want.tolist())


def test_custom_spin_maps_follow_reference():
    want = DATA['lists/sz_blocked_6_0.0_2'].tolist()
    got = st.jw_sz_indices(0.0, 6, 2, up_index=lambda i: i, down_index=lambda i: i + 3)
    assert got == want


def test_sector_argument_errors_match_reference():
    with pytest.raises(ValueError, match='Number of qubits must be even'):
        st.jw_sz_indices(0, 5)
    with pytest.raises(ValueError, match='integer or half-integer'):
        st.jw_sz_indices(0.25, 4)
    with pytest.raises(ValueError, match='incompatible'):
        st.jw_sz_indices(0.5, 4, 2)
    with pytest.raises(ValueError, match='incompatible'):
        sectors.sz_number_sector_tables(4, 1.5, 1)


def test_large_sector_tables_are_consistent_without_enumeration():
    # 4x4 Hubbard, half filling, S_z = 0: 1.66e8 states; spot-check rank/unrank on samples
    tables = sectors.sz_number_sector_tables(32, 0.0, 16)
    assert tables.dim == 12870 ** 2
    rng = numpy.random.default_rng(5)
    up_sites = numpy.array([rng.permutation(16)[:8] for _ in range(200)])
    dn_sites = numpy.array([rng.permutation(16)[:8] for _ in range(200)])
    states = numpy.zeros(200, dtype=numpy.uint64)
    for k in range(200):
        states[k] = sum(1 << (31 - 2 * int(s)) for s in up_sites[k]) + sum(1 << (30 - 2 * int(s)) for s in dn_sites[k])
    assert tables.member(states).all()
    pos = tables.position(states)
    assert pos.min() >= 0 and pos.max() < tables.dim
    # the reference order: lexicographic in (sorted up sites, sorted down sites)
    keys = [(tuple(sorted(u)), tuple(sorted(d))) for u, d in zip(up_sites.tolist(), dn_sites.tolist())]
    assert [keys[i] for i in numpy.argsort(pos)] == sorted(keys)
    first = sum(1 << (31 - 2 * s) for s in range(8)) + sum(1 << (30 - 2 * s) for s in range(8))
    assert tables.position(numpy.array([first], dtype=numpy.uint64))[0] == 0


@pytest.mark.parametrize('name', OPERATORS)
def test_oracle_restriction_matches_reference(name):
    n = int(DATA[name + '/n_qubits'])
    fop = fermion_operator_from(DATA[name + '/fermion_terms'], DATA[name + '/fermion_coeffs'],
                                DATA[name + '/fermion_coeff_is_complex'])
    full = pauli_oracle.jordan_wigner_sparse(fop.terms, n)
    for tag in DATA[name + '/sectors']:
        kind, ne, sz = str(tag).split('_')
        ne = None if ne == 'None' else int(ne)
        idx = (sector_oracle.jw_number_indices(ne, n) if kind == 'number'
               else sector_oracle.jw_sz_indices(float(sz), n, ne))
        got = scipy.sparse.csc_matrix(sector_oracle.restrict_operator(full, idx))
        got.sort_indices()
        key = '{}/{}/f_'.format(name, tag)
        assert numpy.array_equal(got.indptr, DATA[key + 'indptr'])
        assert numpy.array_equal(got.indices, DATA[key + 'indices'])
        assert numpy.array_equal(got.data, DATA[key + 'data'])
        v = DATA['{}/{}/v'.format(name, tag)]
        assert numpy.allclose(got @ v, DATA['{}/{}/Hv'.format(name, tag)], rtol=0, atol=1e-13)


def _np_arguments(key):
    refdet = DATA[key + '/reference_determinant']
    level = int(DATA[key + '/excitation_level'])
    return (str(DATA[key + '/operator']), int(DATA[key + '/num_electrons']),
            bool(DATA[key + '/spin_preserving']), None if refdet.size == 0 else list(refdet),
            None if level < 0 else level)


@pytest.mark.parametrize('key', NP_CASES)
def test_oracle_number_preserving_operator_matches_reference(key):
    name, ne, spin, refdet, level = _np_arguments(key)
    n = int(DATA[name + '/n_qubits'])
    fop = fermion_operator_from(DATA[name + '/fermion_terms'], DATA[name + '/fermion_coeffs'],
                                DATA[name + '/fermion_coeff_is_complex'])
    ordered = ops.normal_ordered(fop)
    got = sector_oracle.get_number_preserving_sparse_operator(ordered.terms, n, ne, spin, refdet, level)
    assert str(got.dtype) == str(DATA[key + '/dtype'])
    assert numpy.array_equal(got.indptr, DATA[key + '/indptr'])
    assert numpy.array_equal(got.indices, DATA[key + '/indices'])
    assert numpy.array_equal(got.data, DATA[key + '/data'])


@pytest.mark.parametrize('key', NP_CASES)
def test_host_basis_enumeration_matches_oracle(key):
    name, ne, spin, refdet, level = _np_arguments(key)
    n = int(DATA[name + '/n_qubits'])
    if refdet is None:
        refdet = [i < ne for i in range(n)]
    if level is None:
        level = ne
    want = [int(numpy.asarray(s, dtype=numpy.int64).dot(1 << numpy.arange(n)[::-1]))
            for s in sector_oracle.iterate_basis(refdet, level, spin)]
    assert st._number_preserving_basis(refdet, level, spin) == want


def test_normal_ordered_known_answers():
    # term_reordering_test.py known answers: 4 3^ -> -3^ 4 ; 1 1^ -> 1 - 1^ 1 ; 2^ 2^ -> 0
    a = ops.normal_ordered(ops.FermionOperator('4 3^'))
    assert a.terms == {((3, 1), (4, 0)): -1.0}
    b = ops.normal_ordered(ops.FermionOperator('1 1^'))
    assert b.terms == {(): 1.0, ((1, 1), (1, 0)): -1.0}
    assert ops.normal_ordered(ops.FermionOperator('2^ 2^')).terms == {}
    c = ops.normal_ordered(ops.FermionOperator('1^ 2^'))
    assert c.terms == {((2, 1), (1, 1)): -1.0}
    with pytest.raises(TypeError):
        ops.normal_ordered(ops.QubitOperator('X0'))


def test_sector_linear_operator_shape_needs_no_gpu():
    from openfermion_b200 import hamiltonians
    from openfermion_b200.linalg import SectorLinearQubitOperator
    hub = hamiltonians.hubbard_jw(2, 2)
    assert SectorLinearQubitOperator(hub, n_electrons=4).shape == (70, 70)
    assert SectorLinearQubitOperator(hub, n_electrons=4, sz_value=0).shape == (36, 36)
    assert SectorLinearQubitOperator(hub, sz_value=0.5).shape == (56, 56)
    with pytest.raises(ValueError):
        SectorLinearQubitOperator(hub)
    with pytest.raises(ValueError):
        SectorLinearQubitOperator(hub, 7, n_electrons=4)
