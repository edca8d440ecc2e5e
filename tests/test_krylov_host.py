"""Host-side pieces of the native Lanczos solver (csrc/krylov.cu) that need no GPU: the Ritz
problem of a sweep, ofb_tridiag_lowest, against scipy's LAPACK path (the routine the Python loop
used: scipy.linalg.eigh_tridiagonal(select='i'))."""
import ctypes

import numpy
import scipy.linalg

from openfermion_b200 import _lib


def _lowest(d, e):
    m = len(d)
    value = ctypes.c_double()
    vector = numpy.zeros(m)
    off = numpy.ascontiguousarray(e if m > 1 else numpy.zeros(1))
    _lib.call('ofb_tridiag_lowest', ctypes.c_int(m), _lib.ptr(numpy.ascontiguousarray(d)), _lib.ptr(off),
              ctypes.byref(value), _lib.ptr(vector))
    return value.value, vector


def test_tridiag_lowest_matches_lapack():
    rng = numpy.random.default_rng(0)
    for trial in range(200):
        m = int(rng.integers(1, 120))
        d = rng.normal(size=m) * rng.choice([1, 10, 0.01])
        e = numpy.abs(rng.normal(size=max(m - 1, 0))) * rng.choice([1, 1e-3, 5])
        if trial % 7 == 0 and m > 3:
            e[m // 2] = 1e-13  # an almost reducible matrix (Lanczos close to an invariant subspace)
        value, vector = _lowest(d, e)
        if m > 1:
            want = scipy.linalg.eigh_tridiagonal(d, e, select='i', select_range=(0, 0))[0][0]
            full = numpy.diag(d) + numpy.diag(e, 1) + numpy.diag(e, -1)
        else:
            want, full = d[0], numpy.array([[d[0]]])
        scale = max(1.0, numpy.abs(full).max())
        assert abs(value - want) <= 1e-14 * scale
        assert abs(numpy.linalg.norm(vector) - 1.0) < 1e-12
        assert numpy.linalg.norm(full @ vector - value * vector) <= 1e-13 * scale


def test_tridiag_lowest_lanczos_like_matrix_with_ghost_copies():
    """A long Lanczos run duplicates converged Ritz values; the lowest one and a vector with a
    small last component (the residual estimate |beta s_m|) must still come out."""
    rng = numpy.random.default_rng(1)
    spectrum = numpy.sort(rng.normal(size=60))
    a = numpy.diag(spectrum)
    v = rng.normal(size=60)
    v /= numpy.linalg.norm(v)
    alphas, betas, prev, beta = [], [], numpy.zeros(60), 0.0
    for _ in range(150):  # no reorthogonalisation: far beyond the 60 exact steps
        w = a @ v - beta * prev
        alpha = v @ w
        w -= alpha * v
        beta = numpy.linalg.norm(w)
        alphas.append(alpha)
        betas.append(beta)
        prev, v = v, w / beta
    value, vector = _lowest(numpy.array(alphas), numpy.array(betas[:-1]))
    assert abs(value - spectrum[0]) < 1e-12
    assert abs(betas[-1] * vector[-1]) < 1e-8
