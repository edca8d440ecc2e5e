"""Shared test helpers: golden-case loading and term-string parsing."""
import os

import numpy

from openfermion_b200.ops import FermionOperator, QubitOperator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _coef(value, keep_complex):
    value = complex(value)
    return value if keep_complex else value.real


def qubit_operator_from(strs, coeffs, is_complex=None):
    op = QubitOperator()
    for k, (s, c) in enumerate(zip(strs, coeffs)):
        keep = True if is_complex is None else bool(is_complex[k])
        key = tuple((int(tok[1:]), tok[0]) for tok in str(s).split())
        op.terms[key] = _coef(c, keep)
    return op


def fermion_operator_from(strs, coeffs, is_complex=None):
    op = FermionOperator()
    for k, (s, c) in enumerate(zip(strs, coeffs)):
        key = tuple((int(tok.rstrip('^')), 1 if tok.endswith('^') else 0) for tok in str(s).split())
        c = complex(c)
        if is_complex is None:
            op.terms[key] = c.real if c.imag == 0 else c
        else:  # keep the reference's coefficient type (it decides the dtype of some outputs)
            op.terms[key] = c if is_complex[k] else c.real
    return op


class Case:
    def __init__(self, data, name):
        self.name = name
        self._d = data

    def __getitem__(self, key):
        return self._d[self.name + '/' + key]

    def has(self, key):
        return (self.name + '/' + key) in self._d.files

    @property
    def n(self):
        return int(self['n_qubits'])


_cache = {}


def golden(file):
    if file not in _cache:
        _cache[file] = numpy.load(os.path.join(GOLDEN, file), allow_pickle=False)
    return _cache[file]


def qubit_cases():
    data = golden('qubit_cases.npz')
    return [Case(data, str(n)) for n in data['names']]


def large_cases():
    data = golden('large_cases.npz')
    return [Case(data, str(n)) for n in data['names']]


def large_state(n, seed):
    """Seeded normalised input of the large golden cases (oracle/make_golden.py:large_state)."""
    rng = numpy.random.default_rng(seed)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / numpy.linalg.norm(v)


def fermion_cases():
    data = golden('fermion_cases.npz')
    return [Case(data, str(n)) for n in data['names']]


def qubit_case_operator(case):
    return qubit_operator_from(case['terms'], case['coeffs'], case['coeff_is_complex'])


def assert_csc_matches(got, indptr, indices, data, rtol=1e-12, signed_zeros=True):
    """CSC parity bar: indptr/indices bit-exact, values within rtol of max |value|."""
    assert got.shape[0] + 1 == len(indptr)
    assert got.indices.dtype == indices.dtype, (got.indices.dtype, indices.dtype)
    assert got.indptr.dtype == indptr.dtype
    assert numpy.array_equal(got.indptr, indptr), 'indptr differs'
    assert numpy.array_equal(got.indices, indices), 'indices differ'
    if len(data):
        scale = numpy.max(numpy.abs(data))
        assert numpy.max(numpy.abs(got.data - data)) <= rtol * scale
        assert got.data.dtype == data.dtype
        if rtol == 0 and signed_zeros:  # "bit for bit" means the bytes: signed zeros included
            assert got.data.tobytes() == numpy.ascontiguousarray(data).tobytes(), 'signed zeros differ'


def close(got, want, rtol=1e-12):
    """|got - want| <= rtol * max|want| (the 1e-12 relative bar of BASELINE.json)."""
    got = numpy.asarray(got)
    want = numpy.asarray(want)
    scale = max(float(numpy.max(numpy.abs(want))) if want.size else 0.0, 1e-300)
    return float(numpy.max(numpy.abs(got - want))) <= rtol * scale if want.size else True


def assert_csc_matches_modulo_residues(got, indptr, indices, data, coeff_abs_sum, rtol=1e-12):
    """Contract for Pauli-route columns holding more than 16 triplets with non-dyadic
    coefficients: scipy's per-column std::sort is unstable there, so
    which floating-point cancellation residues (|v| ~ 1e-16) survive eliminate_zeros depends
    on its internal order.  Entries above 64*eps*sum|c_t| must match in position (bit-exact
    indices) and value (rtol); everything else must be below that floor on both sides."""
    import scipy.sparse
    dim = got.shape[0]
    want = scipy.sparse.csc_matrix((data, indices, indptr), shape=(dim, dim))
    floor = 64 * numpy.finfo(float).eps * coeff_abs_sum

    def strong(m):
        m = m.tocoo()
        keep = numpy.abs(m.data) > floor
        out = scipy.sparse.csc_matrix((m.data[keep], (m.row[keep], m.col[keep])), shape=m.shape)
        out.sort_indices()
        return out

    a, b = strong(got), strong(want)
    assert numpy.array_equal(a.indptr, b.indptr) and numpy.array_equal(a.indices, b.indices)
    assert numpy.max(numpy.abs(a.data - b.data)) <= rtol * numpy.max(numpy.abs(b.data))
    assert abs(got - want).max() <= floor
    assert got.has_sorted_indices and got.indices.dtype == indices.dtype
