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


def fermion_operator_from(strs, coeffs):
    op = FermionOperator()
    for s, c in zip(strs, coeffs):
        key = tuple((int(tok.rstrip('^')), 1 if tok.endswith('^') else 0) for tok in str(s).split())
        c = complex(c)
        op.terms[key] = c.real if c.imag == 0 else c
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


def fermion_cases():
    data = golden('fermion_cases.npz')
    return [Case(data, str(n)) for n in data['names']]


def qubit_case_operator(case):
    return qubit_operator_from(case['terms'], case['coeffs'], case['coeff_is_complex'])


def assert_csc_matches(got, indptr, indices, data, rtol=1e-12):
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


def close(got, want, rtol=1e-12):
    """|got - want| <= rtol * max|want| (the 1e-12 relative bar of BASELINE.json)."""
    got = numpy.asarray(got)
    want = numpy.asarray(want)
    scale = max(float(numpy.max(numpy.abs(want))) if want.size else 0.0, 1e-300)
    return float(numpy.max(numpy.abs(got - want))) <= rtol * scale if want.size else True
