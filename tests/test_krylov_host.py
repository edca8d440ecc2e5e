"""Host-side pieces of the native Lanczos solver (csrc/krylov.cu) that need no GPU: the Ritz
problem of a sweep, ofb_tridiag_lowest, against scipy's LAPACK path (the routine the Python loop
used: scipy.linalg.eigh_tridiagonal(select='i'))."""
import ctypes

import numpy
import pytest
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


# ---- ofb_hermitian_eigh: the projected problem of the native Davidson loop (csrc/davidson.cu) -----
def _eigh(a):
    m = a.shape[0]
    lower = numpy.asfortranarray(numpy.tril(a))  # only the lower triangle is read
    w = numpy.zeros(m)
    z = numpy.zeros((m, m), dtype=numpy.complex128, order='F')
    _lib.call('ofb_hermitian_eigh', ctypes.c_int(m), _lib.ptr(lower), _lib.ptr(w), _lib.ptr(z))
    return w, z


def test_hermitian_eigh_matches_lapack():
    rng = numpy.random.default_rng(7)
    for m in (1, 2, 3, 5, 17, 64, 101):
        for kind in ('complex', 'diagonal', 'degenerate', 'real', 'graded'):
            a = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
            a = a + a.conj().T
            if kind == 'diagonal':
                a = numpy.diag(rng.normal(size=m)).astype(complex)
            elif kind == 'degenerate':
                q, _ = numpy.linalg.qr(a)
                a = q @ numpy.diag(numpy.repeat(rng.normal(size=(m + 2) // 3), 3)[:m]) @ q.conj().T
                a = (a + a.conj().T) / 2
            elif kind == 'real':
                a = a.real.astype(complex)
            elif kind == 'graded':
                s = numpy.diag(10.0 ** -numpy.linspace(0, 8, m))
                a = s @ a @ s
            w, z = _eigh(a)
            scale = max(1.0, numpy.max(numpy.abs(a)))
            assert numpy.all(numpy.diff(w) >= 0)
            assert numpy.max(numpy.abs(w - numpy.linalg.eigvalsh(a))) < 1e-12 * scale * m
            assert numpy.max(numpy.abs(a @ z - z * w)) < 1e-12 * scale * m
            assert numpy.max(numpy.abs(z.conj().T @ z - numpy.eye(m))) < 1e-12 * m


def test_hermitian_eigh_argument_checks():
    with pytest.raises(_lib.OfbError):
        _lib.call('ofb_hermitian_eigh', ctypes.c_int(3), None, None, None)
    _lib.call('ofb_hermitian_eigh', ctypes.c_int(0), None, None, None)
