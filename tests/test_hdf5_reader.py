"""The pure-Python HDF5 reader (openfermion_b200/hdf5.py) on the
reference's bundled molecule files.  The files live in the read-only reference tree, so these
tests run in the build container only (marker needs_reference); the GPU box reads the LiH
integrals from tests/golden/lih_sto3g.npz, which test_lih_file_equals_golden_fixture ties to
the file."""
import glob
import os

import numpy
import pytest

import pauli_oracle
from helpers import golden
from openfermion_b200 import hdf5, hamiltonians
from openfermion_b200.linalg import sparse_tools as st

DATA_DIR = '/root/reference/src/openfermion/testing/data'
FILES = sorted(glob.glob(os.path.join(DATA_DIR, '*.hdf5')))
pytestmark = pytest.mark.needs_reference


def test_all_bundled_files_parse():
    assert len(FILES) == 33
    for path in FILES:
        data = hdf5.load_molecular_data(path)
        assert os.path.basename(path).startswith(data['name'])
        assert data['n_qubits'] == 2 * data['n_orbitals']
        n = data['n_orbitals']
        assert data['one_body_integrals'].shape == (n, n)
        assert data['two_body_integrals'].shape == (n, n, n, n)
        assert data['one_body_integrals'].dtype == numpy.float64
        assert len(data['geometry']) == data['n_atoms']
        assert data['fci_energy'] < data['hf_energy']


def test_lih_scalars_and_structure():
    path = os.path.join(DATA_DIR, 'H1-Li1_sto-3g_singlet_1.45.hdf5')
    with hdf5.File(path) as f:
        assert 'fci_energy' in f and 'nonexistent' not in f
        assert sorted(f['geometry'].keys()) == ['atoms', 'positions']
        assert f['two_body_integrals'].shape == (6, 6, 6, 6)
        assert f['general_calculations_keys'][...].dtype == bool  # "missing" marker of the reference
        with pytest.raises(KeyError):
            f['geometry/none']
    data = hdf5.load_molecular_data(path[:-5])  # the reference passes the name without .hdf5
    # chem/molecular_data_test.py / lih_integration_test.py known values
    assert data['nuclear_repulsion'] == 1.0948493970827586
    assert float(data['hf_energy']) == -7.8625677857178955
    assert float(data['fci_energy']) == -7.8809823148256966
    assert float(data['mp2_energy']) == -7.874480485350199
    assert float(data['cisd_energy']) == -7.880970958694532
    assert float(data['ccsd_energy']) == -7.880973454314876
    assert data['geometry'] == [('Li', [0.0, 0.0, 0.0]), ('H', [0.0, 0.0, 1.45])]
    assert (data['multiplicity'], data['charge'], data['n_electrons'], data['n_atoms']) == (1, 0, 4, 2)
    assert data['basis'] == 'sto-3g' and data['description'] == '1.45'
    assert list(data['protons']) == [1, 3]
    # overlap of orthonormal canonical orbitals: C^T S C = 1
    c, s = data['canonical_orbitals'], data['overlap_integrals']
    assert numpy.allclose(c.T @ s @ c, numpy.eye(6), atol=1e-10)
    # the one-particle density matrix of the FCI state traces to the electron count
    assert abs(numpy.trace(data['fci_one_rdm']) - 4.0) < 1e-8


def test_lih_file_equals_golden_fixture():
    g = golden('lih_sto3g.npz')
    data = hdf5.load_molecular_data(os.path.join(DATA_DIR, 'H1-Li1_sto-3g_singlet_1.45.hdf5'))
    assert numpy.array_equal(data['one_body_integrals'], g['one_body_integrals'])
    assert numpy.array_equal(data['two_body_integrals'], g['two_body_integrals'])
    assert data['nuclear_repulsion'] == float(g['nuclear_repulsion'])


@pytest.mark.parametrize('path', [p for p in FILES if '/H2_' in p][::4] + [p for p in FILES if '6-31g' in p],
                         ids=os.path.basename)
def test_integrals_reproduce_the_stored_fci_energy(path):
    # file -> integrals -> InteractionOperator terms -> JW matrix (CPU oracle) -> lowest eigenvalue
    ham = hdf5.load_molecular_hamiltonian(path)
    fop = st._tensor_fermion_terms(ham)
    matrix = pauli_oracle.jordan_wigner_sparse(fop.terms, ham.n_qubits).toarray()
    e0 = numpy.linalg.eigvalsh(matrix)[0]
    data = hdf5.load_molecular_data(path)
    assert abs(e0 - float(data['fci_energy'])) < 1e-7
    # Hartree-Fock determinant energy = stored hf_energy
    hf_index = sum(1 << (ham.n_qubits - 1 - k) for k in range(data['n_electrons']))
    assert abs(matrix[hf_index, hf_index].real - float(data['hf_energy'])) < 1e-7


def test_unsupported_files_fail_loudly(tmp_path):
    bad = tmp_path / 'x.hdf5'
    bad.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(hdf5.Hdf5FormatError):
        hdf5.File(str(bad))
    raw = bytearray(open(FILES[0], 'rb').read())
    raw[8] = 2  # pretend superblock version 2
    newer = tmp_path / 'y.hdf5'
    newer.write_bytes(bytes(raw))
    with pytest.raises(hdf5.Hdf5FormatError, match='superblock version 2'):
        hdf5.File(str(newer))


def test_reduced_density_matrix_files():
    # testing/{tqdm,phdm}_*.hdf5: one contiguous float64 dataset each
    base = os.path.dirname(DATA_DIR)
    for name, n in (('H2_sto-3g_singlet_1.4', 4), ('H2_6-31g_singlet_0.75', 8),
                    ('H1-Li1_sto-3g_singlet_1.45', 12)):
        tq = hdf5.File(os.path.join(base, 'tqdm_{}.hdf5'.format(name)))['tqdm'][...]
        ph = hdf5.File(os.path.join(base, 'phdm_{}.hdf5'.format(name)))['phdm'][...]
        assert tq.shape == ph.shape == (n, n, n, n) and tq.dtype == numpy.float64
        # two-hole RDM: antisymmetric in its first index pair, trace fixed by the hole count
        assert numpy.abs(tq + tq.transpose(1, 0, 2, 3)).max() < 1e-12
        n_electrons = {4: 2, 8: 2, 12: 4}[n]
        holes = n - n_electrons
        assert abs(numpy.einsum('pqqp', tq) - holes * (holes - 1)) < 1e-8
