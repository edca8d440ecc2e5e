"""The Lanczos solver behind the C ABI (csrc/krylov.cu, ofb_lanczos_lowest and the ofb_kws_*
primitives) against the host-driven loop it replaces (OFB_LANCZOS_HOST=1), dense eigenvalues,
and -- for the fused pieces -- the separate kernels: the apply launch that also leaves
conj(v) . (H v), and the one-pass update."""
import ctypes

import numpy
import pytest
import scipy.sparse

import helpers
import pauli_oracle as po
from openfermion_b200 import _lib, hamiltonians
from openfermion_b200.linalg import (LinearQubitOperator, SectorLinearQubitOperator, get_gap,
                                     get_ground_state, get_sparse_operator, krylov)
from openfermion_b200.ops import FermionOperator

pytestmark = pytest.mark.gpu


OPERATORS = {  # built inside the test: construction of some of them needs the device
    'plan_hubbard_n16': lambda: LinearQubitOperator(hamiltonians.hubbard_jw(2, 4), 16),
    'plan_molecular_n12': lambda: LinearQubitOperator(hamiltonians.molecular_jw(6, seed=1234), 12),
    # real coefficients on random strings: a Hermitian Pauli sum with Y factors
    'plan_random_strings_n10': lambda: LinearQubitOperator(hamiltonians.random_pauli_sum(10, 120, 7, False), 10),
    'sector_hubbard_3x4_half': lambda: SectorLinearQubitOperator(hamiltonians.hubbard_jw(3, 4), 24,
                                                                 n_electrons=12, sz_value=0.0),
    'csr_molecular_n12': lambda: get_sparse_operator(hamiltonians.molecular_jw(6, seed=1234), 12),
}


@pytest.mark.parametrize('name', list(OPERATORS))
def test_native_solver_matches_host_loop(name, monkeypatch):
    op = OPERATORS[name]()
    stats_native, stats_host = {}, {}
    dev = krylov.as_device_operator(op)
    e_native, vec, res = krylov.lanczos_lowest(dev, stats=stats_native)
    assert res <= 1e-8 * max(1.0, abs(e_native))
    monkeypatch.setenv('OFB_LANCZOS_HOST', '1')
    e_host, _, _ = krylov.lanczos_lowest(dev, stats=stats_host)
    assert abs(e_native - e_host) <= 1e-10 * max(1.0, abs(e_host))
    # the returned vector is a unit eigenvector: Rayleigh quotient and residual recomputed here
    n = dev.dim
    hv = _lib.DeviceBuffer(n * 16)
    dev.apply(vec.ptr, hv.ptr)
    assert abs(krylov.nrm2(n, vec.ptr) - 1.0) < 1e-12
    assert abs(krylov.dotc(n, vec.ptr, hv.ptr).real - e_native) <= 1e-10 * max(1.0, abs(e_native))
    assert stats_native['matvecs'] <= 1.2 * stats_host['matvecs'] + 8


def test_native_solver_small_and_degenerate_cases():
    # dimension 2 and 4 (fewer Krylov steps than the default), a diagonal operator (breakdown
    # after one step from an eigenvector-like start), and an explicit initial guess
    for terms, n in (({((0, 'Z'),): 1.0, ((0, 'X'),): 0.5}, 1),
                     ({((0, 'Z'), (1, 'Z')): 1.0, ((0, 'X'),): 0.3, ((1, 'Y'),): -0.2}, 2),
                     ({((0, 'Z'),): 1.0, ((2, 'Z'),): 2.0, (): 0.5}, 3)):
        from openfermion_b200.ops import QubitOperator
        op = QubitOperator()
        op.terms.update(terms)
        dense = po.qubit_operator_sparse(op.terms, n).toarray()
        want = numpy.linalg.eigvalsh(dense)[0]
        energy, state = get_ground_state(LinearQubitOperator(op, n))
        assert abs(energy - want) < 1e-10
        assert numpy.linalg.norm(dense @ state - energy * state) < 1e-8
        guess = numpy.ones(1 << n, dtype=complex)
        energy, _ = get_ground_state(LinearQubitOperator(op, n), guess)
        assert abs(energy - want) < 1e-10


def test_get_gap_with_locking_matches_dense():
    op = hamiltonians.molecular_jw(4, seed=9)
    mat = get_sparse_operator(op, 8)
    values = numpy.linalg.eigvalsh(mat.toarray())
    assert abs(get_gap(mat) - (values[1] - values[0])) < 1e-8


def test_fused_apply_dot_and_step_equal_the_separate_kernels():
    op = hamiltonians.molecular_jw(7, seed=3)
    n = 14
    dim = 1 << n
    lin = LinearQubitOperator(op, n)
    plan = lin.plan
    u = helpers.large_state(n, 5)
    up = helpers.large_state(n, 6)
    d_u, d_up = _lib.to_device(u), _lib.to_device(up)
    d_w, d_ref = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    ws = ctypes.c_void_p()
    _lib.call('ofb_kws_create', ctypes.c_int64(dim), ctypes.c_int(8), ctypes.c_int(0), ctypes.byref(ws))
    try:
        _lib.call('ofb_kws_reset', ws, ctypes.c_double(1.0), None)
        _lib.call('ofb_apply_groups_dot', plan._apply_handle, ctypes.c_void_p(d_u.ptr), ctypes.c_void_p(d_w.ptr),
                  ctypes.c_int(n), ctypes.c_uint64(0), ctypes.c_int64(0), ctypes.c_int64(plan.n_groups),
                  ctypes.c_int(0), ctypes.c_void_p(d_u.ptr), ws, None)
        plan.apply_device(d_u.ptr, d_ref.ptr)
        w = krylov.download(dim, d_w.ptr)
        assert numpy.array_equal(w, krylov.download(dim, d_ref.ptr))
        want_dot = numpy.vdot(u, w)
        _lib.call('ofb_kws_alpha', ws, ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0), None)
        _lib.call('ofb_kws_step', ws, ctypes.c_void_p(d_w.ptr), ctypes.c_void_p(d_u.ptr), None, None,
                  ctypes.c_int(0), None)
        _lib.call('ofb_kws_beta', ws, ctypes.c_int(0), None)
        alphas, betas = numpy.zeros(1), numpy.zeros(2)
        flag = ctypes.c_int()
        _lib.call('ofb_kws_fetch', ws, ctypes.c_int(1), _lib.ptr(alphas), _lib.ptr(betas), ctypes.byref(flag), None)
        assert flag.value == 0 and betas[0] == 1.0
        assert abs(alphas[0] - want_dot.real) <= 1e-12 * abs(want_dot)
        w1 = w - alphas[0] * u
        got = krylov.download(dim, d_w.ptr)
        assert helpers.close(got, w1, 1e-13)
        assert abs(betas[1] - numpy.linalg.norm(w1)) <= 1e-12 * betas[1]
        # second step with a previous vector: w <- H(w1)/b1 - (a1/b1) w1 - (b1/b0) u  (u_prev = u)
        d_w2 = _lib.DeviceBuffer(dim * 16)
        _lib.call('ofb_apply_groups_dot', plan._apply_handle, ctypes.c_void_p(d_w.ptr), ctypes.c_void_p(d_w2.ptr),
                  ctypes.c_int(n), ctypes.c_uint64(0), ctypes.c_int64(0), ctypes.c_int64(plan.n_groups),
                  ctypes.c_int(0), ctypes.c_void_p(d_w.ptr), ws, None)
        _lib.call('ofb_kws_alpha', ws, ctypes.c_int(1), ctypes.c_int(0), ctypes.c_double(0.0), None)
        _lib.call('ofb_kws_step', ws, ctypes.c_void_p(d_w2.ptr), ctypes.c_void_p(d_w.ptr), ctypes.c_void_p(d_u.ptr),
                  None, ctypes.c_int(0), None)
        _lib.call('ofb_kws_beta', ws, ctypes.c_int(1), None)
        alphas, betas = numpy.zeros(2), numpy.zeros(3)
        _lib.call('ofb_kws_fetch', ws, ctypes.c_int(2), _lib.ptr(alphas), _lib.ptr(betas), ctypes.byref(flag), None)
        v1 = w1 / betas[1]
        hv1 = po.matvec(op.terms, n, v1)
        a1 = numpy.vdot(v1, hv1).real
        want = hv1 - a1 * v1 - betas[1] * u
        assert abs(alphas[1] - a1) <= 1e-11 * max(1.0, abs(a1))
        assert helpers.close(krylov.download(dim, d_w2.ptr), want, 1e-11)
    finally:
        _lib.call('ofb_kws_destroy', ws)


def test_fermion_operator_ground_state_at_particle_number():
    """a FermionOperator handed to jw_get_ground_state_at_particle_number."""
    from openfermion_b200.linalg import jordan_wigner_sparse, jw_get_ground_state_at_particle_number
    op = FermionOperator()
    for p, q in ((0, 2), (2, 4), (1, 3), (3, 5), (0, 4), (1, 5)):
        op += FermionOperator(((p, 1), (q, 0)), -1.0) + FermionOperator(((q, 1), (p, 0)), -1.0)
    for s in (0, 2, 4):
        op += FermionOperator(((s, 1), (s, 0), (s + 1, 1), (s + 1, 0)), 4.0)
    full = jordan_wigner_sparse(op, 6).toarray()
    for particles in (2, 3):
        index = [j for j in range(64) if bin(j).count('1') == particles]
        want = numpy.linalg.eigvalsh(full[numpy.ix_(index, index)])[0]
        energy, state = jw_get_ground_state_at_particle_number(op, particles)
        assert abs(energy - want) < 1e-10
        assert state.shape == (64,) and numpy.linalg.norm(full @ state - energy * state) < 1e-8
        sector_energy, sector_state = jw_get_ground_state_at_particle_number(op, particles, expand=False)
        assert abs(sector_energy - want) < 1e-10 and sector_state.shape == (len(index),)
