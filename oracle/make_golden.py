"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from the REAL reference.

Run in the build container only (needs /root/reference through oracle/ref_shim.py):
    python oracle/make_golden.py
Every case stores the operator (term strings + coefficients), the seeded inputs and the
reference's own outputs, and asserts on the way that the oracle restatement
(oracle/pauli_oracle.py) reproduces them, so the GPU box -- which has neither the
reference nor this script's inputs -- can check both the oracle and the CUDA path.
"""
import hashlib
import os
import sys
import zlib

import numpy
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402

ref_shim.load()
import pauli_oracle as po  # noqa: E402
from openfermion.hamiltonians.hubbard import fermi_hubbard  # noqa: E402
from openfermion.linalg import (LinearQubitOperator, get_ground_state,  # noqa: E402
                                get_linear_qubit_operator_diagonal, get_sparse_operator,
                                jordan_wigner_sparse, qubit_operator_sparse, expectation)
from openfermion.ops.operators import QubitOperator  # noqa: E402
from openfermion.ops.representations import InteractionOperator  # noqa: E402
from openfermion.chem.molecular_data import spinorb_from_spatial  # noqa: E402
from openfermion.testing.testing_utils import (random_interaction_operator,  # noqa: E402
                                               random_qubit_operator)
from openfermion.transforms.opconversions import (bravyi_kitaev, get_fermion_operator,  # noqa: E402
                                                  jordan_wigner)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
LIH = '/root/reference/src/openfermion/testing/data/H1-Li1_sto-3g_singlet_1.45.hdf5'


def qubit_term_str(term):
    return ' '.join('{}{}'.format(a, q) for q, a in term)


def fermion_term_str(term):
    return ' '.join('{}{}'.format(q, '^' if a else '') for q, a in term)


def pack_terms(terms, to_str):
    strs = numpy.array([to_str(t) for t in terms], dtype=str)
    coeffs = numpy.array([complex(c) for c in terms.values()], dtype=numpy.complex128)
    is_complex = numpy.array([isinstance(c, (complex, numpy.complexfloating)) for c in terms.values()])
    return strs, coeffs, is_complex


def same_csc(a, b):
    return (a.shape == b.shape and a.indices.dtype == b.indices.dtype
            and numpy.array_equal(a.indptr, b.indptr) and numpy.array_equal(a.indices, b.indices)
            and numpy.array_equal(a.data, b.data))


def sha(a):
    return hashlib.sha256(numpy.ascontiguousarray(a).tobytes()).hexdigest()


def qubit_cases():
    rng = numpy.random.default_rng(20261019)
    cases = {}
    ops = {}
    for seed in range(8):
        n = 3 + seed
        ops['random_q%d_s%d' % (n, seed)] = (random_qubit_operator(n, 24, n, seed=seed), n)
    ops['hubbard_2x2_jw'] = (jordan_wigner(fermi_hubbard(2, 2, 1.0, 4.0)), 8)
    ops['hubbard_2x2_mu_h_jw'] = (jordan_wigner(fermi_hubbard(2, 2, 1.0, 4.0, 0.3, 0.2)), 8)
    ops['hubbard_2x2_bk'] = (bravyi_kitaev(fermi_hubbard(2, 2, 1.0, 4.0)), 8)
    ops['hubbard_2x3_jw'] = (jordan_wigner(fermi_hubbard(2, 3, 1.0, 4.0)), 12)
    mol = random_interaction_operator(3, expand_spin=True, real=True, seed=5)
    ops['molecular_6_jw'] = (jordan_wigner(mol), 6)
    ops['molecular_6_bk'] = (bravyi_kitaev(get_fermion_operator(mol)), 6)
    ops['molecular_8_jw'] = (jordan_wigner(random_interaction_operator(4, expand_spin=True, real=True, seed=1234)), 8)
    cplx = random_interaction_operator(3, expand_spin=False, real=False, seed=11)
    ops['molecular_3_complex_jw'] = (jordan_wigner(cplx), 3)
    ops['padded_n7'] = (QubitOperator('X0 Y1 Z3', 0.5 - 0.25j) + QubitOperator('Z1 Z2', 1.5), 7)
    ops['identity_only'] = (QubitOperator((), 2.5), 4)
    names = []
    for name, (op, n) in ops.items():
        v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
        strs, coeffs, is_complex = pack_terms(op.terms, qubit_term_str)
        ref_mv = LinearQubitOperator(op, n) * v
        assert numpy.array_equal(ref_mv, po.matvec(op.terms, n, v)), name
        ref_csc = qubit_operator_sparse(op, n)
        assert same_csc(ref_csc, po.qubit_operator_sparse(op.terms, n)), name
        ref_exp = expectation(LinearQubitOperator(op, n), v)
        cases[name + '/terms'] = strs
        cases[name + '/coeffs'] = coeffs
        cases[name + '/coeff_is_complex'] = is_complex
        cases[name + '/n_qubits'] = numpy.array(n)
        cases[name + '/vec'] = v
        cases[name + '/matvec'] = ref_mv
        cases[name + '/expectation'] = numpy.array(ref_exp)
        cases[name + '/indptr'] = ref_csc.indptr
        cases[name + '/indices'] = ref_csc.indices
        cases[name + '/data'] = ref_csc.data
        if not is_complex.any():
            ref_diag = get_linear_qubit_operator_diagonal(op, n)
            assert numpy.array_equal(ref_diag, po.diagonal(op.terms, n)), name
            cases[name + '/diagonal'] = ref_diag
        names.append(name)
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'qubit_cases.npz'), **cases)
    print('qubit_cases:', len(names), 'cases')


def fermion_cases():
    cases = {}
    ops = {
        'hubbard_2x2': (fermi_hubbard(2, 2, 1.0, 4.0), 8),
        'hubbard_2x2_mu_h': (fermi_hubbard(2, 2, 1.0, 4.0, 0.3, 0.2), 8),
        'hubbard_2x3_open': (fermi_hubbard(2, 3, 1.0, 2.0, periodic=False), 12),
        'molecular_6': (get_fermion_operator(random_interaction_operator(3, expand_spin=True, real=True, seed=5)), 6),
        'molecular_8_dense': (get_fermion_operator(random_interaction_operator(8, expand_spin=False, real=True, seed=3)), 8),
        'molecular_4_complex': (get_fermion_operator(random_interaction_operator(4, expand_spin=False, real=False, seed=9)), 4),
    }
    names = []
    for name, (op, n) in ops.items():
        strs, coeffs, _ = pack_terms(op.terms, fermion_term_str)
        ref_csc = jordan_wigner_sparse(op, n)
        assert same_csc(ref_csc, po.jordan_wigner_sparse(op.terms, n)), name
        cases[name + '/terms'] = strs
        cases[name + '/coeffs'] = coeffs
        cases[name + '/n_qubits'] = numpy.array(n)
        cases[name + '/indptr'] = ref_csc.indptr
        cases[name + '/indices'] = ref_csc.indices
        cases[name + '/data'] = ref_csc.data
        names.append(name)
    # DiagonalCoulombHamiltonian through get_sparse_operator (sparse_tools_test.py:1876-1890)
    from openfermion.testing.testing_utils import random_diagonal_coulomb_hamiltonian
    dch = random_diagonal_coulomb_hamiltonian(5, real=False, seed=3)
    ref_csc = get_sparse_operator(dch)
    cases['dch/one_body'] = dch.one_body
    cases['dch/two_body'] = dch.two_body
    cases['dch/constant'] = numpy.array(dch.constant)
    cases['dch/indptr'] = ref_csc.indptr
    cases['dch/indices'] = ref_csc.indices
    cases['dch/data'] = ref_csc.data
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'fermion_cases.npz'), **cases)
    print('fermion_cases:', len(names), 'cases')


def lih_case():
    """Config 1: LiH sto-3g from the bundled HDF5.  No h5py here: the two gzip-chunked
    float64 datasets are read at their byte offsets (one_body_integrals 6x6 at 2775,
    two_body_integrals 6^4 at 22704; SURVEY.md section 7 hard part 5) and validated by
    reproducing the file's stored fci_energy through the reference."""
    raw = open(LIH, 'rb').read()
    one = numpy.frombuffer(zlib.decompressobj().decompress(raw[2775:]), dtype='<f8').reshape(6, 6).copy()
    two = numpy.frombuffer(zlib.decompressobj().decompress(raw[22704:]), dtype='<f8').reshape(6, 6, 6, 6).copy()
    nuclear_repulsion = 1.0948493970827586
    fci_energy = -7.8809823148256966
    hf_energy = -7.8625677857178955
    one_so, two_so = spinorb_from_spatial(one, two)
    ham = InteractionOperator(nuclear_repulsion, one_so, 0.5 * two_so)
    csc = get_sparse_operator(ham)
    e0, psi = get_ground_state(csc)
    assert abs(e0 - fci_energy) < 1e-7, (e0, fci_energy)
    assert same_csc(csc, po.jordan_wigner_sparse(get_fermion_operator(ham).terms, 12))
    qham = jordan_wigner(ham)
    hf_state = numpy.zeros(2**12)
    hf_state[sum(2 ** (11 - i) for i in range(4))] = 1.0
    e_hf = expectation(csc, hf_state)
    assert abs(e_hf - hf_energy) < 1e-7
    strs, coeffs, is_complex = pack_terms(qham.terms, qubit_term_str)
    numpy.savez_compressed(
        os.path.join(GOLDEN, 'lih_sto3g.npz'),
        one_body_integrals=one, two_body_integrals=two, nuclear_repulsion=nuclear_repulsion,
        fci_energy=fci_energy, hf_energy=hf_energy, reference_e0=e0, reference_hf_expectation=e_hf,
        nnz=csc.nnz, indptr=csc.indptr, indices_sha256=sha(csc.indices), indices_dtype=str(csc.indices.dtype),
        data_abs_sum=numpy.abs(csc.data).sum(), data_sum=csc.data.sum(),
        data_head=csc.data[:512], indices_head=csc.indices[:512],
        qubit_terms=strs, qubit_coeffs=coeffs, n_qubit_terms=len(strs),
    )
    print('lih: nnz', csc.nnz, 'E0', e0, 'T', len(strs))


def sector_cases():
    """Symmetry-sector functions: index lists, restricted
    operators, get_number_preserving_sparse_operator and sector ground states of the real
    reference; the oracle restatement (oracle/sector_oracle.py) and the repo's own
    normal_ordered are asserted against the reference on the way."""
    import sector_oracle as so
    from openfermion.linalg import sparse_tools as rst
    from openfermion.transforms.opconversions.term_reordering import normal_ordered
    from openfermion_b200 import ops as our_ops
    cases = {}
    # 1. index lists
    lists = []
    for n, ne in ((1, 0), (1, 1), (5, 2), (6, 3), (7, 0), (7, 7), (9, 4)):
        ref = rst.jw_number_indices(ne, n)
        assert ref == so.jw_number_indices(ne, n)
        key = 'number_{}_{}'.format(n, ne)
        cases['lists/' + key] = numpy.array(ref, dtype=numpy.int64)
        lists.append(key)
    for n, sz, ne in ((2, 0.5, None), (4, 0.0, None), (6, 0.5, None), (6, -1.0, None), (8, 0.0, 4),
                      (6, 0.5, 3), (6, -1.5, 3), (8, 1.0, 4), (10, -0.5, 5), (8, 2.0, None)):
        ref = rst.jw_sz_indices(sz, n, ne)
        assert ref == so.jw_sz_indices(sz, n, ne)
        key = 'sz_{}_{}_{}'.format(n, sz, ne)
        cases['lists/' + key] = numpy.array(ref, dtype=numpy.int64)
        lists.append(key)
    # custom spin maps (first half up, second half down)
    ref = rst.jw_sz_indices(0.0, 6, 2, up_index=lambda i: i, down_index=lambda i: i + 3)
    assert ref == so.jw_sz_indices(0.0, 6, 2, up_index=lambda i: i, down_index=lambda i: i + 3)
    cases['lists/sz_blocked_6_0.0_2'] = numpy.array(ref, dtype=numpy.int64)
    cases['list_names'] = numpy.array(lists, dtype=str)

    # 2. restricted operators + 5. sector matvec + 4. ground states
    rng = numpy.random.default_rng(20261020)
    fops = {
        'hubbard_2x2': (fermi_hubbard(2, 2, 1.0, 4.0), 8),
        'hubbard_2x2_mu_h': (fermi_hubbard(2, 2, 1.0, 4.0, 0.3, 0.2), 8),
        'molecular_6': (get_fermion_operator(random_interaction_operator(3, expand_spin=True, real=True, seed=5)), 6),
        'molecular_8': (get_fermion_operator(random_interaction_operator(4, expand_spin=True, real=True, seed=11)), 8),
        'molecular_4_complex': (get_fermion_operator(random_interaction_operator(4, expand_spin=False, real=False, seed=9)), 4),
    }
    sector_list = {
        'hubbard_2x2': [('number', 4, None), ('number', 3, None), ('sz', 4, 0.0), ('sz', None, 0.0), ('sz', 3, 0.5)],
        'hubbard_2x2_mu_h': [('number', 5, None), ('sz', None, -1.0)],
        'molecular_6': [('number', 3, None), ('sz', 3, 0.5), ('sz', None, -0.5), ('sz', 2, 0.0)],
        'molecular_8': [('number', 4, None), ('sz', 4, 0.0), ('sz', None, 1.0)],
        'molecular_4_complex': [('number', 2, None), ('sz', 2, 0.0)],
    }
    names = []
    for name, (fop, n) in fops.items():
        qop = jordan_wigner(fop)
        fs, fc, fic = pack_terms(fop.terms, fermion_term_str)
        qs, qc, qic = pack_terms(qop.terms, qubit_term_str)
        cases[name + '/n_qubits'] = numpy.array(n)
        cases[name + '/fermion_terms'] = fs
        cases[name + '/fermion_coeffs'] = fc
        cases[name + '/fermion_coeff_is_complex'] = fic
        cases[name + '/qubit_terms'] = qs
        cases[name + '/qubit_coeffs'] = qc
        cases[name + '/qubit_coeff_is_complex'] = qic
        bs, bc, _ = pack_terms(bravyi_kitaev(fop).terms, qubit_term_str)
        cases[name + '/bk_terms'] = bs
        cases[name + '/bk_coeffs'] = bc
        full_f = get_sparse_operator(fop, n)
        full_q = get_sparse_operator(qop, n)
        tags = []
        for kind, ne, sz in sector_list[name]:
            tag = '{}_{}_{}'.format(kind, ne, sz)
            if kind == 'number':
                res_f = rst.jw_number_restrict_operator(full_f, ne, n)
                res_q = rst.jw_number_restrict_operator(full_q, ne, n)
                idx = so.jw_number_indices(ne, n)
            else:
                res_f = rst.jw_sz_restrict_operator(full_f, sz, ne, n)
                res_q = rst.jw_sz_restrict_operator(full_q, sz, ne, n)
                idx = so.jw_sz_indices(sz, n, ne)
            assert same_csc(scipy.sparse.csc_matrix(res_f), scipy.sparse.csc_matrix(so.restrict_operator(full_f, idx)))
            for route, res in (('f', res_f), ('q', res_q)):
                res = scipy.sparse.csc_matrix(res)
                res.sort_indices()
                cases['{}/{}/{}_indptr'.format(name, tag, route)] = res.indptr
                cases['{}/{}/{}_indices'.format(name, tag, route)] = res.indices
                cases['{}/{}/{}_data'.format(name, tag, route)] = res.data
            dim = res_f.shape[0]
            v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
            cases['{}/{}/v'.format(name, tag)] = v
            cases['{}/{}/Hv'.format(name, tag)] = res_f @ v
            cases['{}/{}/expectation'.format(name, tag)] = numpy.array(expectation(scipy.sparse.csc_matrix(res_f), v))
            cases['{}/{}/diagonal'.format(name, tag)] = res_f.diagonal()
            if dim > 2:
                cases['{}/{}/e0'.format(name, tag)] = numpy.array(
                    numpy.linalg.eigvalsh(res_f.toarray())[0])
            tags.append(tag)
        cases[name + '/sectors'] = numpy.array(tags, dtype=str)
        names.append(name)
    # jw_get_ground_state_at_particle_number (sparse_tools_test.py style known answers)
    for name, ne in (('hubbard_2x2', 2), ('hubbard_2x2', 4), ('molecular_6', 3), ('molecular_6', 0)):
        fop, n = fops[name]
        e, state = rst.jw_get_ground_state_at_particle_number(get_sparse_operator(fop, n), ne)
        e2, _ = so.ground_state_at_particle_number(get_sparse_operator(fop, n), ne)
        assert abs(e - e2) < 1e-10
        cases['{}/gs_number_{}/energy'.format(name, ne)] = numpy.array(e)
        cases['{}/gs_number_{}/support'.format(name, ne)] = numpy.nonzero(numpy.abs(state) > 1e-12)[0]
    cases['names'] = numpy.array(names, dtype=str)

    # 3. get_number_preserving_sparse_operator
    np_cases = []
    specs = [
        ('hubbard_2x2', 4, False, None, None),
        ('hubbard_2x2', 4, True, None, None),
        ('hubbard_2x2', 3, True, [True, False, True, True, False, False, False, False], None),
        ('molecular_6', 3, False, None, 2),
        ('molecular_6', 2, True, [False, True, True, False, False, False], 1),
        ('molecular_8', 4, True, None, 2),
        ('molecular_8', 4, False, [True, True, False, False, True, True, False, False], None),
        ('molecular_4_complex', 2, False, None, None),
    ]
    for k, (name, ne, spin, refdet, level) in enumerate(specs):
        fop, n = fops[name]
        ref = rst.get_number_preserving_sparse_operator(fop, n, ne, spin_preserving=spin,
                                                         reference_determinant=refdet,
                                                         excitation_level=level)
        # the repo's normal ordering must be bit-identical to the reference's (term order too)
        ordered_ref = normal_ordered(fop)
        mine = our_ops.FermionOperator()
        mine.terms = dict(fop.terms)
        ordered_mine = our_ops.normal_ordered(mine)
        assert list(ordered_ref.terms.items()) == list(ordered_mine.terms.items()), name
        again = so.get_number_preserving_sparse_operator(ordered_ref.terms, n, ne, spin, refdet, level)
        assert same_csc(ref, again), (name, k)
        key = 'np{}'.format(k)
        cases[key + '/operator'] = numpy.array(name)
        cases[key + '/num_electrons'] = numpy.array(ne)
        cases[key + '/spin_preserving'] = numpy.array(spin)
        cases[key + '/reference_determinant'] = numpy.array(refdet if refdet is not None else [], dtype=bool)
        cases[key + '/excitation_level'] = numpy.array(-1 if level is None else level)
        cases[key + '/indptr'] = ref.indptr
        cases[key + '/indices'] = ref.indices
        cases[key + '/data'] = ref.data
        cases[key + '/dtype'] = numpy.array(str(ref.dtype))
        np_cases.append(key)
    cases['np_names'] = numpy.array(np_cases, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'sector_cases.npz'), **cases)
    print('sector_cases:', len(lists), 'lists,', len(names), 'operators,', len(np_cases), 'number-preserving cases')


def operator_file_cases():
    """Plain-text operator files written by the reference's save_operator
    (utils/operator_utils.py:309-360), for openfermion_b200/operator_files.py."""
    from openfermion.utils.operator_utils import save_operator, load_operator
    fop = fermi_hubbard(2, 2, 1.0, 4.0, 0.3, 0.2)
    qop = jordan_wigner(fop) + QubitOperator('X0 Y3', 0.5 - 0.25j)
    for name, op in (('operator_file_fermion', fop), ('operator_file_qubit', qop)):
        path = os.path.join(GOLDEN, name + '.data')
        if os.path.exists(path):
            os.remove(path)
        save_operator(op, name, GOLDEN, plain_text=True)
        assert load_operator(name, GOLDEN, plain_text=True) == op
    print('operator files written')


def ladder_cases():
    """Single ladder-operator matrices of the reference (sparse_tools.py:49-73) and the
    Kronecker helpers they are built with (:39-46)."""
    from openfermion.linalg import jordan_wigner_ladder_sparse, kronecker_operators, wrapped_kronecker
    cases = {}
    names = []
    for n, mode, kind in ((1, 0, 1), (1, 0, 0), (3, 0, 1), (3, 2, 0), (5, 2, 1), (6, 5, 0), (9, 4, 1), (11, 0, 0)):
        ref = jordan_wigner_ladder_sparse(n, mode, kind)
        assert same_csc(ref, po.jordan_wigner_sparse({((mode, kind),): 1.0}, n)), (n, mode, kind)
        name = 'n%d_m%d_%s' % (n, mode, 'raise' if kind else 'lower')
        cases[name + '/args'] = numpy.array([n, mode, kind])
        cases[name + '/indptr'] = ref.indptr
        cases[name + '/indices'] = ref.indices
        cases[name + '/data'] = ref.data
        cases[name + '/index_dtype'] = numpy.array(str(ref.indices.dtype))
        names.append(name)
    rng = numpy.random.default_rng(11)
    a = scipy.sparse.random(3, 4, 0.5, format='csc', random_state=rng, dtype=float) * (1 + 0.5j)
    b = scipy.sparse.random(2, 2, 0.8, format='csc', random_state=rng, dtype=float)
    c = scipy.sparse.identity(3, dtype=complex, format='csc')
    for tag, mat in (('ab', wrapped_kronecker(a, b)), ('abc', kronecker_operators([a, b, c]))):
        cases['kron_' + tag + '/indptr'] = mat.indptr
        cases['kron_' + tag + '/indices'] = mat.indices
        cases['kron_' + tag + '/data'] = mat.data
        cases['kron_' + tag + '/shape'] = numpy.array(mat.shape)
    for tag, mat in (('a', a), ('b', b), ('c', c)):
        cases['kron_in_' + tag] = mat.toarray()
    # gate-building helpers (sparse_tools.py:516-554); theta = 0 stores explicit zeros
    from openfermion.linalg import (jw_sparse_givens_rotation,
                                    jw_sparse_particle_hole_transformation_last_mode)
    givens = [(0, 1, 0.3, 0.2, 3), (1, 2, 0.0, 0.0, 4), (0, 1, numpy.pi / 2, 0.1, 2), (4, 5, 1.1, -0.4, 6)]
    cases['givens_args'] = numpy.array(givens, dtype=float)
    for k, (i, j, theta, phi, n) in enumerate(givens):
        mat = jw_sparse_givens_rotation(i, j, theta, phi, n)
        cases['givens%d/indptr' % k] = mat.indptr
        cases['givens%d/indices' % k] = mat.indices
        cases['givens%d/data' % k] = mat.data
    for n in (1, 3, 5):
        mat = jw_sparse_particle_hole_transformation_last_mode(n)
        cases['particle_hole%d/indptr' % n] = mat.indptr
        cases['particle_hole%d/indices' % n] = mat.indices
        cases['particle_hole%d/data' % n] = mat.data
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'ladder_cases.npz'), **cases)
    print('ladder_cases:', len(names), 'cases')


def large_state(n, seed):
    """Seeded normalised input of the large cases; tests/helpers.py:large_state is the same code."""
    rng = numpy.random.default_rng(seed)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / numpy.linalg.norm(v)


def large_cases():
    """The parity ladder of SURVEY.md section 8d above the sizes whose full outputs fit a fixture:
    the real reference's LinearQubitOperator._matvec, expectation and diagonal at 14-20 qubits --
    the sizes at which the tiled CUDA kernel (not the small-state kernel) runs.  Stored: the
    operator, the input's seed and sha256, 2 048 sampled output amplitudes, <v|Hv>, |Hv| and
    sum(Hv); the diagonal likewise.  The numpy and C oracles are checked on the way."""
    import time
    import c_oracle
    from openfermion_b200 import compiler
    cases = {}
    names = []
    ops = {
        'hubbard_2x4_jw_n16': (jordan_wigner(fermi_hubbard(2, 4, 1.0, 4.0)), 16),
        'hubbard_2x5_jw_n20': (jordan_wigner(fermi_hubbard(2, 5, 1.0, 4.0)), 20),
        'hubbard_2x4_bk_n16': (bravyi_kitaev(fermi_hubbard(2, 4, 1.0, 4.0, 0.3, 0.2)), 16),
        'molecular_7_jw_n14': (jordan_wigner(random_interaction_operator(7, expand_spin=True, real=True, seed=1234)), 14),
        'molecular_4_complex_bk_n8_padded_n15': (bravyi_kitaev(get_fermion_operator(
            random_interaction_operator(8, expand_spin=False, real=False, seed=21))), 15),
        'molecular_8_jw_n16': (jordan_wigner(random_interaction_operator(8, expand_spin=True, real=True, seed=1234)), 16),
        # BASELINE.json config 3 at full size: one H|psi> of the real reference (minutes)
        'hubbard_3x4_jw_n24': (jordan_wigner(fermi_hubbard(3, 4, 1.0, 4.0)), 24),
    }
    rng = numpy.random.default_rng(99)
    for k, (name, (op, n)) in enumerate(ops.items()):
        seed = 1000 + k
        v = large_state(n, seed)
        t0 = time.time()
        ref_mv = LinearQubitOperator(op, n) * v
        t_ref = time.time() - t0
        strs, coeffs, is_complex = pack_terms(op.terms, qubit_term_str)
        x, z, c = compiler.compile_qubit_terms(op.terms, n)
        got_c = c_oracle.matvec(n, x, z, c, v)
        scale = numpy.max(numpy.abs(ref_mv))
        assert numpy.max(numpy.abs(got_c - ref_mv)) <= 1e-13 * scale, name
        if len(op.terms) <= 200 and n <= 20:  # the numpy port walks the same ~20 passes per term
            assert numpy.array_equal(po.matvec(op.terms, n, v), ref_mv), name
        sample = numpy.sort(rng.choice(2**n, size=2048, replace=False))
        cases[name + '/terms'] = strs
        cases[name + '/coeffs'] = coeffs
        cases[name + '/coeff_is_complex'] = is_complex
        cases[name + '/n_qubits'] = numpy.array(n)
        cases[name + '/seed'] = numpy.array(seed)
        cases[name + '/vec_sha256'] = numpy.array(sha(v))
        cases[name + '/sample_index'] = sample
        cases[name + '/sample_matvec'] = ref_mv[sample]
        cases[name + '/expectation'] = numpy.array(numpy.vdot(v, ref_mv))
        cases[name + '/norm'] = numpy.array(numpy.linalg.norm(ref_mv))
        cases[name + '/sum'] = numpy.array(numpy.sum(ref_mv))
        cases[name + '/max_abs'] = numpy.array(scale)
        if not is_complex.any() and n <= 16:
            ref_diag = get_linear_qubit_operator_diagonal(op, n)
            cases[name + '/sample_diagonal'] = ref_diag[sample]
            cases[name + '/diagonal_sum'] = numpy.array(numpy.sum(ref_diag))
        names.append(name)
        print('  %s: T = %d, reference matvec %.1f s' % (name, len(op.terms), t_ref))
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'large_cases.npz'), **cases)
    print('large_cases:', len(names), 'cases')


def csc_ladder_cases():
    """CSC parity ladder (the real reference up to n = 14 molecular,
    n = 16 Hubbard): the reference's own matrices by sha256 of indptr / indices / data, plus
    sampled entries so that a mismatch can be located.  Pauli route and fermion route."""
    import time
    cases = {}
    names = []
    mol6 = random_interaction_operator(6, expand_spin=True, real=True, seed=7)
    mol7 = random_interaction_operator(7, expand_spin=True, real=True, seed=1234)
    todo = {
        'pauli/hubbard_2x4_jw_n16': ('pauli', jordan_wigner(fermi_hubbard(2, 4, 1.0, 4.0)), 16),
        'pauli/hubbard_2x4_bk_n16': ('pauli', bravyi_kitaev(fermi_hubbard(2, 4, 1.0, 4.0)), 16),
        'fermion/hubbard_2x4_mu_h_n16': ('fermion', fermi_hubbard(2, 4, 1.0, 4.0, 0.3, 0.2), 16),
        'pauli/molecular_6_jw_n12': ('pauli', jordan_wigner(mol6), 12),
        'fermion/molecular_6_n12': ('fermion', get_fermion_operator(mol6), 12),
        'pauli/molecular_7_jw_n14': ('pauli', jordan_wigner(mol7), 14),
        'fermion/molecular_7_n14': ('fermion', get_fermion_operator(mol7), 14),
    }
    rng = numpy.random.default_rng(5)
    for name, (route, op, n) in todo.items():
        t0 = time.time()
        if route == 'pauli':
            ref = qubit_operator_sparse(op, n)
            strs, coeffs, is_complex = pack_terms(op.terms, qubit_term_str)
        else:
            ref = jordan_wigner_sparse(op, n)
            strs, coeffs, is_complex = pack_terms(op.terms, fermion_term_str)
        dt = time.time() - t0
        if n <= 12:  # the scipy-based oracle needs the same triplet memory as the reference
            mine = po.qubit_operator_sparse(op.terms, n) if route == 'pauli' else po.jordan_wigner_sparse(op.terms, n)
            assert same_csc(ref, mine), name
        key = name.replace('/', '__')
        sample = numpy.sort(rng.choice(ref.nnz, size=min(4096, ref.nnz), replace=False))
        cases[key + '/route'] = numpy.array(route)
        cases[key + '/terms'] = strs
        cases[key + '/coeffs'] = coeffs
        cases[key + '/coeff_is_complex'] = is_complex
        cases[key + '/n_qubits'] = numpy.array(n)
        cases[key + '/nnz'] = numpy.array(ref.nnz)
        cases[key + '/index_dtype'] = numpy.array(str(ref.indices.dtype))
        cases[key + '/indptr_sha256'] = numpy.array(sha(ref.indptr))
        cases[key + '/indices_sha256'] = numpy.array(sha(ref.indices))
        cases[key + '/data_sha256'] = numpy.array(sha(ref.data))
        cases[key + '/sample_position'] = sample
        cases[key + '/sample_indices'] = ref.indices[sample]
        cases[key + '/sample_data'] = ref.data[sample]
        cases[key + '/colcount_head'] = numpy.diff(ref.indptr)[:256]
        cases[key + '/data_abs_sum'] = numpy.array(numpy.abs(ref.data).sum())
        names.append(key)
        print('  %s: T = %d, nnz = %d, reference %.1f s' % (name, len(op.terms), ref.nnz, dt))
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'csc_ladder_cases.npz'), **cases)
    print('csc_ladder_cases:', len(names), 'cases')


def energy_ladder_cases():
    """Eigenvalue ladder (ground state by the CSC route and matrix-free up
    to n = 16, Davidson at n = 12): the reference's own get_ground_state / get_gap /
    jw_get_ground_state_at_particle_number / QubitDavidson results.  Spectra are degenerate:
    only energies are stored."""
    import time
    from openfermion.linalg import (QubitDavidson, get_gap, jw_get_ground_state_at_particle_number,
                                    variance)
    cases = {}
    names = []
    ops = {
        'hubbard_2x4_n16': (jordan_wigner(fermi_hubbard(2, 4, 1.0, 4.0)), 16, (2, 4, 8)),
        'hubbard_2x3_mu_h_n12': (jordan_wigner(fermi_hubbard(2, 3, 1.0, 4.0, 0.3, 0.2)), 12, (3, 6)),
        'molecular_6_n12': (jordan_wigner(random_interaction_operator(6, expand_spin=True, real=True, seed=7)), 12, (2, 6)),
        'molecular_7_n14': (jordan_wigner(random_interaction_operator(7, expand_spin=True, real=True, seed=1234)), 14, (4,)),
    }
    for name, (op, n, numbers) in ops.items():
        t0 = time.time()
        mat = get_sparse_operator(op, n)
        e0, state = get_ground_state(mat)
        gap = get_gap(mat)
        e0_linear = None
        if n <= 12:  # matrix-free ARPACK through the reference's LinearQubitOperator: slow
            e0_linear, _ = get_ground_state(LinearQubitOperator(op, n))
        strs, coeffs, is_complex = pack_terms(op.terms, qubit_term_str)
        cases[name + '/terms'] = strs
        cases[name + '/coeffs'] = coeffs
        cases[name + '/coeff_is_complex'] = is_complex
        cases[name + '/n_qubits'] = numpy.array(n)
        cases[name + '/e0'] = numpy.array(e0)
        cases[name + '/gap'] = numpy.array(gap)
        cases[name + '/variance_of_ground_state'] = numpy.array(variance(mat, state))
        if e0_linear is not None:
            cases[name + '/e0_matrix_free'] = numpy.array(e0_linear)
        cases[name + '/particle_numbers'] = numpy.array(numbers)
        cases[name + '/e0_at_particle_number'] = numpy.array(
            [jw_get_ground_state_at_particle_number(mat, k)[0] for k in numbers])
        if n == 12:  # the reference's Davidson test size (davidson_test.py)
            import copy
            real_op = copy.deepcopy(op)
            real_op.compress()  # complex-typed coefficients make the reference's diagonal raise
            numpy.random.seed(7)
            success, values, _ = QubitDavidson(real_op, n).get_lowest_n(2, max_iterations=60)
            cases[name + '/davidson_success'] = numpy.array(bool(success))
            cases[name + '/davidson_lowest2'] = numpy.array(values)
        names.append(name)
        print('  %s: E0 = %.12f gap = %.3e  (%.1f s)' % (name, e0, gap, time.time() - t0))
    cases['names'] = numpy.array(names, dtype=str)
    numpy.savez_compressed(os.path.join(GOLDEN, 'energy_ladder_cases.npz'), **cases)
    print('energy_ladder_cases:', len(names), 'cases')


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    which = sys.argv[1:] or ['qubit', 'fermion', 'lih', 'sector', 'files', 'ladder']
    if 'ladder' in which:
        ladder_cases()
    if 'large' in which:  # minutes of reference time: not part of the default set
        large_cases()
    if 'energy_ladder' in which:
        energy_ladder_cases()
    if 'csc_ladder' in which:  # gigabytes of triplets in the reference: not part of the default set
        csc_ladder_cases()
    if 'qubit' in which:
        qubit_cases()
    if 'fermion' in which:
        fermion_cases()
    if 'lih' in which:
        lih_case()
    if 'sector' in which:
        sector_cases()
    if 'files' in which:
        operator_file_cases()
