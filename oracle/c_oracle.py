"""TEST INFRASTRUCTURE ONLY -- ctypes loader of the C oracle (oracle/pauli_oracle.c)."""
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, 'pauli_oracle.c')
LIB = os.path.join(_HERE, '_build', 'libpauli_oracle.so')


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-shared', '-fPIC',
                           '-o', LIB, SRC])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def matvec(n_qubits, x_mask, z_mask, coeff, vec):
    """Reference-order matvec from index-convention masks (see pauli_oracle.c)."""
    x = numpy.ascontiguousarray(x_mask, dtype=numpy.uint64)
    z = numpy.ascontiguousarray(z_mask, dtype=numpy.uint64)
    c = numpy.ascontiguousarray(coeff, dtype=numpy.complex128)
    v = numpy.ascontiguousarray(vec, dtype=numpy.complex128)
    out = numpy.empty_like(v)
    p = ctypes.c_void_p
    _load().oracle_matvec(ctypes.c_int(n_qubits), ctypes.c_int64(len(c)), p(x.ctypes.data),
                          p(z.ctypes.data), p(c.ctypes.data), p(v.ctypes.data), p(out.ctypes.data))
    return out


def diagonal(n_qubits, x_mask, z_mask, coeff):
    x = numpy.ascontiguousarray(x_mask, dtype=numpy.uint64)
    z = numpy.ascontiguousarray(z_mask, dtype=numpy.uint64)
    c = numpy.ascontiguousarray(coeff, dtype=numpy.complex128)
    out = numpy.empty(1 << n_qubits, dtype=numpy.float64)
    p = ctypes.c_void_p
    _load().oracle_diagonal(ctypes.c_int(n_qubits), ctypes.c_int64(len(c)), p(x.ctypes.data),
                            p(z.ctypes.data), p(c.ctypes.data), p(out.ctypes.data))
    return out
