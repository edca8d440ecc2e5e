"""Davidson's method behind the C ABI (csrc/davidson.cu, ofb_davidson_lowest_n) against the
host-driven loop it replaces (OFB_DAVIDSON_HOST=1, the loop that restates
linalg/davidson.py:99-339 of the reference) and against dense eigenvalues."""
import ctypes
import logging
import warnings

import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import _lib, hamiltonians
from openfermion_b200.linalg import (DavidsonOptions, LinearQubitOperator, QubitDavidson, SparseDavidson,
                                     get_linear_qubit_operator_diagonal, get_sparse_operator, krylov)

pytestmark = pytest.mark.gpu


def _dense_eigs(op, n):
    return numpy.linalg.eigvalsh(po.qubit_operator_sparse(op.terms, n).toarray())


@pytest.mark.parametrize('case', ['hubbard_2x3', 'molecular_n10', 'random_strings_n9'])
@pytest.mark.parametrize('n_lowest', [1, 3])
def test_native_equals_host_loop_and_dense(case, n_lowest, monkeypatch):
    op, n = {'hubbard_2x3': (hamiltonians.hubbard_jw(2, 3), 12),
             'molecular_n10': (hamiltonians.molecular_jw(5, seed=77), 10),
             'random_strings_n9': (hamiltonians.random_pauli_sum(9, 80, 5, False), 9)}[case]
    rng = numpy.random.default_rng(5)
    guess = rng.random((1 << n, n_lowest)) + 1j * rng.random((1 << n, n_lowest))
    options = DavidsonOptions(max_subspace=40, max_iterations=400, eps=1e-7)
    dav = QubitDavidson(op, n, options=options)
    ok, values, vectors = dav.get_lowest_n(n_lowest, guess)
    assert ok and dav.last_iterations > 0
    lin = LinearQubitOperator(op, n)
    assert numpy.max(numpy.abs(lin.matmat(vectors) - vectors * values)) < 1e-5
    assert numpy.max(numpy.abs(vectors.conj().T @ vectors - numpy.eye(n_lowest))) < 1e-10
    # Davidson with the diagonal preconditioner finds eigenvalues, not necessarily the lowest
    # ones of a degenerate spectrum in order: every value is an eigenvalue, the first the lowest
    dense = _dense_eigs(op, n)
    for v in values:
        assert numpy.min(numpy.abs(dense - v)) < 1e-8
    monkeypatch.setenv('OFB_DAVIDSON_HOST', '1')
    ok_h, values_h, _ = QubitDavidson(op, n, options=options).get_lowest_n(n_lowest, guess)
    assert ok_h
    assert numpy.allclose(values, values_h, rtol=0, atol=1e-8)


def test_native_trajectory_equals_host_loop(monkeypatch):
    # same Ritz values after a fixed number of iterations (not converged): the native loop
    # follows the host loop step by step, restarts included (max_subspace 6)
    op, n = hamiltonians.molecular_jw(6, seed=3), 12
    guess = numpy.random.default_rng(2).random((1 << n, 2))
    for max_subspace, iterations in ((40, 7), (6, 11)):
        options = DavidsonOptions(max_subspace=max_subspace, eps=1e-9)
        ok, values, _ = QubitDavidson(op, n, options=options).get_lowest_n(2, guess, max_iterations=iterations)
        monkeypatch.setenv('OFB_DAVIDSON_HOST', '1')
        ok_h, values_h, _ = QubitDavidson(op, n, options=options).get_lowest_n(2, guess, max_iterations=iterations)
        monkeypatch.delenv('OFB_DAVIDSON_HOST')
        assert ok == ok_h and not ok
        assert numpy.allclose(values, values_h, rtol=0, atol=1e-9 * max(1.0, numpy.max(numpy.abs(values_h))))


def test_native_sparse_davidson_and_real_only():
    op, n = hamiltonians.hubbard_jw(2, 2), 8
    mat = get_sparse_operator(op, n)
    dense = numpy.linalg.eigvalsh(mat.toarray())
    for real_only in (False, True):
        options = DavidsonOptions(max_subspace=30, eps=1e-8, real_only=real_only)
        numpy.random.seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter('error')
            ok, values, vectors = SparseDavidson(mat, options=options).get_lowest_n(2)
        assert ok
        for v in values:
            assert numpy.min(numpy.abs(dense - v)) < 1e-8
        assert abs(values[0] - dense[0]) < 1e-8
        if real_only:
            assert numpy.allclose(vectors.imag, 0.0)


def test_info_logging_keeps_the_reference_log_line(caplog):
    op, n = hamiltonians.hubbard_jw(2, 2), 8
    with caplog.at_level(logging.INFO):
        ok, _, _ = QubitDavidson(op, n).get_lowest_n(1, numpy.ones((1 << n, 1)) + numpy.arange(1 << n)[:, None])
    assert ok and any('Eigenvalues for iteration' in r.getMessage() for r in caplog.records)


def test_c_abi_call_with_seeded_random_start_and_argument_checks():
    op, n = hamiltonians.molecular_jw(5, seed=1), 10
    dim = 1 << n
    lin = LinearQubitOperator(op, n)
    dev = krylov.as_device_operator(lin)
    diag = _lib.to_device(numpy.ascontiguousarray(get_linear_qubit_operator_diagonal(op, n).real))
    out = _lib.DeviceBuffer(dim * 16 * 2)
    values = numpy.zeros(2)
    ok, its, flags = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    err = ctypes.c_double()
    handle = dev.linop()
    try:
        def call(n_lowest, max_subspace, eps=1e-7):
            _lib.call('ofb_davidson_lowest_n', handle, ctypes.c_void_p(diag.ptr), ctypes.c_int(n_lowest),
                      ctypes.c_int(0), None, ctypes.c_int(max_subspace), ctypes.c_int(300),
                      ctypes.c_double(eps), ctypes.c_int(0), ctypes.c_uint64(42), _lib.ptr(values),
                      ctypes.c_void_p(out.ptr), ctypes.byref(ok), ctypes.byref(its), ctypes.byref(err),
                      ctypes.byref(flags), None)
        call(2, 30)
        assert ok.value == 1 and err.value < 1e-7 and its.value > 0
        dense = _dense_eigs(op, n)
        assert abs(values[0] - dense[0]) < 1e-8 and numpy.min(numpy.abs(dense - values[1])) < 1e-8
        vectors = krylov.download(2 * dim, out.ptr).reshape(2, dim).T
        assert numpy.max(numpy.abs(lin.matmat(vectors) - vectors * values)) < 1e-5
        for bad in ((0, 30), (30, 30), (1, 2)):  # n_lowest outside [1, max_subspace), max_subspace <= 2
            with pytest.raises(_lib.OfbError):
                call(*bad)
        with pytest.raises(_lib.OfbError):
            call(1, 30, eps=0.0)
    finally:
        _lib.call('ofb_linop_destroy', handle)
        diag.free()
        out.free()
