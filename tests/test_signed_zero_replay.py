"""Signed zeros of the reference's matrix entries (include/ofb200.h, ofb_pauli_entry_zero_sign;
csrc/ofb_internal.cuh sz_replay): the 3-bit replay of scipy.sparse.kron's multiplication chain
against the chain itself -- the oracle's pauli_term_triplets, which is bit-identical to the real
reference on every golden case -- for random strings, every column, and every kind of
coefficient, including the ones whose zeros are already negative on entry."""
import ctypes
import itertools

import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import _lib

COEFFS = [1.5, -1.5, 2, -3, complex(1.5, 0.0), complex(-1.5, 0.0), complex(-1.5, -0.0), complex(1.5, -0.0),
          1.5j, -1.5j, complex(0.0, -1.5), complex(-0.0, 1.5), complex(0.25, -0.75), -(1.5 + 0j)]


def _zero_sign(n, x, z, coefficient, column):
    c = complex(coefficient)
    has, neg = ctypes.c_int(), ctypes.c_int()
    _lib.call('ofb_pauli_entry_zero_sign', ctypes.c_int(n), ctypes.c_uint64(x), ctypes.c_uint64(z),
              ctypes.c_double(c.real), ctypes.c_double(c.imag), ctypes.c_uint64(column),
              ctypes.byref(has), ctypes.byref(neg))
    return has.value, neg.value


def _check(term, coefficient, n):
    data, _, cols = po.pauli_term_triplets(term, coefficient, n)
    x, z, _ = po.term_masks(term, n)
    for value, column in zip(data, cols.astype(numpy.int64)):
        has, neg = _zero_sign(n, x, z, coefficient, int(column))
        zero_re, zero_im = value.real == 0.0, value.imag == 0.0
        assert has == int(zero_re != zero_im), (term, coefficient, column)
        if has:
            part = value.real if zero_re else value.imag
            assert neg == int(numpy.signbit(part)), (term, coefficient, column, value)


def test_all_strings_on_three_qubits():
    for actions in itertools.product('IXYZ', repeat=3):
        term = tuple((q, a) for q, a in enumerate(actions) if a != 'I')
        for coefficient in COEFFS:
            _check(term, coefficient, 3)


@pytest.mark.parametrize('n', [1, 2, 5, 7])
def test_random_strings(n):
    rng = numpy.random.default_rng(n)
    for _ in range(60):
        actions = rng.choice(list('IXYZ'), size=n, p=[0.4, 0.2, 0.2, 0.2])
        term = tuple((q, str(a)) for q, a in enumerate(actions) if a != 'I')
        _check(term, COEFFS[int(rng.integers(len(COEFFS)))], n)


def test_transition_tables_come_from_numpy_multiplication():
    """The four tables in csrc/ofb_internal.cuh, regenerated from numpy's complex product."""
    factors = {'E': 1 + 0j, 'N': -1 + 0j, 'Y0': 1j, 'Y1': -1j}
    want = {'E': 0xdac680, 'N': 0x937052, 'Y0': 0x212de4, 'Y1': 0x48193e}

    def make(state):
        nz = -1.5 if (state >> 1) & 1 else 1.5
        zero = -0.0 if state & 1 else 0.0
        return complex(nz, zero) if not state >> 2 else complex(zero, nz)

    def decode(v):
        if v.imag == 0:
            return (int(v.real < 0) << 1) | int(numpy.signbit(v.imag))
        return 4 | (int(v.imag < 0) << 1) | int(numpy.signbit(v.real))
    for name, factor in factors.items():
        table = 0
        for state in range(8):
            product = (numpy.array([make(state)]) * numpy.array([factor]))[0]
            table |= decode(product) << (3 * state)
        assert table == want[name], name
