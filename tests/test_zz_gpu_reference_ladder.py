"""CSC parity ladder against the REAL reference at the largest sizes it can build here
(molecular n <= 14, Hubbard n <= 16; tests/golden/csc_ladder_cases.npz,
written by oracle/make_golden.py csc_ladder): nnz, index width, indptr and indices bit-exact by
sha256; values within 1e-12 of max |value| on sampled entries and -- in the scipy-order mode
that is the default at these sizes -- bit-exact by sha256 as well.  Runs last in the suite."""
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
    assert _sha(mat.data) == str(case['data_sha256']), 'values differ from the reference in the last bits'
