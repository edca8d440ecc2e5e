"""BASELINE.json config 2 at FULL size on the GPU: the 20-qubit synthetic molecular-type
Hamiltonian (dyadic integrals, T = 22 351, 7.4e8 non-zeros) materialised as CSC
(qubit_operator_sparse, linalg/sparse_tools.py:136-194) and checked against the CPU oracle.

The reference cannot build this matrix (~750 GB of triplets), so the oracle is evaluated on
column blocks (oracle/pauli_oracle.py:qubit_operator_sparse_columns: the same triplets -> coo ->
csc -> eliminate_zeros pipeline restricted to the block; exact because the coefficients are
dyadic): six blocks of 1 024 consecutive columns incl. the first and the last, plus 64 seeded
single columns.  Structure (indptr differences, row indices) and values must be IDENTICAL."""
import hashlib

import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import hamiltonians
from openfermion_b200.linalg import LinearQubitOperator, expectation, qubit_operator_sparse

pytestmark = pytest.mark.gpu

N = 20


@pytest.fixture(scope='module')
def config2():
    op = hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16)
    return op, qubit_operator_sparse(op, N)


def _block_equal(mat, block, columns):
    lo, hi = mat.indptr[columns[0]], mat.indptr[columns[-1] + 1]
    contiguous = numpy.array_equal(columns, numpy.arange(columns[0], columns[-1] + 1))
    if contiguous:
        assert numpy.array_equal(mat.indptr[columns[0]:columns[-1] + 2] - lo, block.indptr)
        assert numpy.array_equal(mat.indices[lo:hi], block.indices)
        assert numpy.array_equal(mat.data[lo:hi], block.data)
        return hashlib.sha256(mat.indices[lo:hi].tobytes()).hexdigest()
    for k, c in enumerate(columns):
        a, b = mat.indptr[c], mat.indptr[c + 1]
        p, q = block.indptr[k], block.indptr[k + 1]
        assert numpy.array_equal(mat.indices[a:b], block.indices[p:q]), c
        assert numpy.array_equal(mat.data[a:b], block.data[p:q]), c
    return None


def test_config2_csc_against_oracle_blocks(config2):
    op, mat = config2
    dim = 1 << N
    assert mat.shape == (dim, dim) and mat.dtype == numpy.complex128
    assert mat.indices.dtype == numpy.int32 and mat.indptr.dtype == numpy.int32  # nnz < 2^31
    assert mat.indptr[0] == 0 and mat.indptr[-1] == mat.nnz == len(mat.indices) == len(mat.data)
    assert mat.nnz > 7e8
    rng = numpy.random.default_rng(20)
    starts = [0, dim - 1024] + [int(s) * 1024 for s in rng.integers(1, dim // 1024 - 1, size=4)]
    terms = op.terms
    for start in starts:
        columns = numpy.arange(start, start + 1024)
        block = po.qubit_operator_sparse_columns(terms, N, columns)
        digest = _block_equal(mat, block, columns)
        assert digest == hashlib.sha256(block.indices.tobytes()).hexdigest()
    singles = numpy.unique(rng.integers(0, dim, size=64))
    _block_equal(mat, po.qubit_operator_sparse_columns(terms, N, singles), singles)


def test_config2_rows_ascend_and_matrix_is_hermitian_on_samples(config2):
    _, mat = config2
    rng = numpy.random.default_rng(21)
    for c in rng.integers(0, 1 << N, size=200):
        a, b = mat.indptr[c], mat.indptr[c + 1]
        rows = mat.indices[a:b]
        assert numpy.all(numpy.diff(rows) > 0)
        # H[r, c] == conj(H[c, r]) for a few entries of the column
        for k in (0, (b - a) // 2, b - a - 1):
            r = int(rows[k])
            p, q = mat.indptr[r], mat.indptr[r + 1]
            at = numpy.searchsorted(mat.indices[p:q], c)
            assert mat.indices[p + at] == c and mat.data[p + at] == numpy.conj(mat.data[a + k])


def test_config2_expectation_matches_the_materialised_matrix(config2):
    op, mat = config2
    rng = numpy.random.default_rng(0)
    psi = rng.normal(size=1 << N) + 1j * rng.normal(size=1 << N)
    psi /= numpy.linalg.norm(psi)
    want = numpy.vdot(psi, mat @ psi)  # host product with the CSC the oracle blocks certify
    got = expectation(LinearQubitOperator(op, N), psi)
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
    assert abs(expectation(mat, psi) - want) <= 1e-12 * max(1.0, abs(want))
