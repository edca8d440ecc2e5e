"""Device-resident Krylov machinery: operator adapters, vector helpers and the Lanczos
ground-state loop behind get_ground_state / get_gap.

The reference delegates to ARPACK (sparse_tools.py:578: eigsh(k=1, which='SA', v0,
maxiter=1e7), ncv=20 complex Arnoldi).  Twenty basis vectors do not fit at 28-32 qubits
(SURVEY.md section 7 hard part 4), so the B200 path is a three-term Lanczos recurrence
with all vectors resident in HBM, Ritz values of the small tridiagonal matrix on the
host, a second pass that rebuilds the Ritz vector, and restarts from that vector until
the true residual ||H y - theta y|| is below tolerance.  The returned energy is the
Rayleigh quotient of the returned vector.
"""
import ctypes
import os

import numpy
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

from openfermion_b200 import _lib
from openfermion_b200._lib import c_void_p, c_int, c_int64, c_double

C128 = 16


# ----------------------------------------------------------------------------
# thin wrappers over the K5 entry points
# ----------------------------------------------------------------------------
def _ri(value):
    value = complex(value)
    return (c_double * 2)(value.real, value.imag)


def dotc(n, x, y, stream=None):
    out = (c_double * 2)()
    _lib.call('ofb_vec_dotc', c_int64(n), c_void_p(x), c_void_p(y), out, c_void_p(stream))
    return complex(out[0], out[1])


def nrm2(n, x, stream=None):
    out = c_double()
    _lib.call('ofb_vec_nrm2', c_int64(n), c_void_p(x), ctypes.byref(out), c_void_p(stream))
    return out.value


def maxabs(n, x, stream=None):
    out = c_double()
    _lib.call('ofb_vec_maxabs', c_int64(n), c_void_p(x), ctypes.byref(out), c_void_p(stream))
    return out.value


def axpy(n, alpha, x, y, stream=None):
    _lib.call('ofb_vec_axpy', c_int64(n), _ri(alpha), c_void_p(x), c_void_p(y), c_void_p(stream))


def scal(n, alpha, x, stream=None):
    _lib.call('ofb_vec_scal', c_int64(n), _ri(alpha), c_void_p(x), c_void_p(stream))


def multi_dotc(n, ncols, V, ld, w, stream=None):
    out = (c_double * (2 * max(ncols, 1)))()
    if ncols:
        _lib.call('ofb_vec_multi_dotc', c_int64(n), c_int(ncols), c_void_p(V), c_int64(ld),
                  c_void_p(w), out, c_void_p(stream))
    return numpy.array([complex(out[2 * k], out[2 * k + 1]) for k in range(ncols)])


def multi_axpy(n, coeffs, V, ld, beta, w, stream=None):
    coeffs = numpy.ascontiguousarray(coeffs, dtype=numpy.complex128)
    _lib.call('ofb_vec_multi_axpy', c_int64(n), c_int(len(coeffs)),
              coeffs.ctypes.data_as(ctypes.POINTER(c_double)), c_void_p(V), c_int64(ld), _ri(beta),
              c_void_p(w), c_void_p(stream))


def lanczos_update(n, alpha, beta, v, v_prev, w, stream=None):
    _lib.call('ofb_vec_lanczos_update', c_int64(n), c_double(alpha), c_double(beta), c_void_p(v),
              c_void_p(v_prev), c_void_p(w), c_void_p(stream))


def copy(n, src, dst, stream=None):
    _lib.call('ofb_memcpy_d2d', c_void_p(dst), c_void_p(src), ctypes.c_size_t(n * C128),
              c_void_p(stream))


def upload(host_vec, dst, stream=None):
    host_vec = numpy.ascontiguousarray(host_vec, dtype=numpy.complex128)
    _lib.copy_h2d(dst, host_vec, stream)


def download(n, src, stream=None):
    out = numpy.empty(n, dtype=numpy.complex128)
    _lib.copy_d2h(out, src, stream)
    return out


def fill_random(n, seed, real_only, dst, stream=None):
    _lib.call('ofb_vec_fill_random', c_int64(n), ctypes.c_uint64(seed), c_int(1 if real_only else 0),
              c_void_p(dst), c_void_p(stream))


# ----------------------------------------------------------------------------
# operator adapters: everything is applied on device pointers
# ----------------------------------------------------------------------------
class PlanOperator:
    """A compiled Pauli plan (K1)."""

    def __init__(self, plan):
        self.plan = plan
        self.dim = plan.dim

    def apply(self, d_in, d_out, stream=None):
        self.plan.apply_device(d_in, d_out, 1, self.dim, stream)

    def apply_block(self, d_in, d_out, ncols, ld, stream=None):
        """ncols columns `ld` elements apart in one launch (K1b: the per-term work is shared)."""
        self.plan.apply_device(d_in, d_out, ncols, ld, stream)

    def linop(self):
        """New ofb_linop handle (caller destroys it) for the solvers behind the C ABI."""
        handle = c_void_p()
        _lib.call('ofb_linop_from_plan', self.plan._apply_handle, ctypes.byref(handle))
        return handle


class SectorOperator:
    """A compiled Pauli plan restricted to a symmetry sector (K7, csrc/sector.cu)."""

    def __init__(self, plan, sector):
        self.plan = plan
        self.sector = sector
        self.dim = sector.dim

    def apply(self, d_in, d_out, stream=None):
        self.sector.apply_device(self.plan, d_in, d_out, 1, self.dim, stream)

    def linop(self):
        handle = c_void_p()
        _lib.call('ofb_linop_from_sector', self.plan.handle, self.sector.handle, ctypes.byref(handle))
        return handle


class CsrOperator:
    """An already materialised scipy sparse matrix, uploaded once as CSR."""

    def __init__(self, matrix):
        _lib.require_device()
        csr = scipy.sparse.csr_matrix(matrix)
        csr.sum_duplicates()
        self.dim = csr.shape[0]
        self.index_bits = 64 if csr.indices.dtype == numpy.int64 else 32
        idx = numpy.int64 if self.index_bits == 64 else numpy.int32
        self._indptr = _lib.to_device(csr.indptr.astype(idx))
        self._indices = _lib.to_device(csr.indices.astype(idx))
        self._data = _lib.to_device(csr.data.astype(numpy.complex128))

    def apply(self, d_in, d_out, stream=None):
        _lib.call('ofb_csr_matvec', c_int64(self.dim), c_void_p(self._indptr.ptr),
                  c_void_p(self._indices.ptr), c_void_p(self._data.ptr), c_int(self.index_bits),
                  c_void_p(d_in), c_void_p(d_out), c_void_p(stream))

    def linop(self):
        handle = c_void_p()
        _lib.call('ofb_linop_from_csr', c_int64(self.dim), c_void_p(self._indptr.ptr),
                  c_void_p(self._indices.ptr), c_void_p(self._data.ptr), c_int(self.index_bits),
                  c_int(0), ctypes.byref(handle))
        return handle


class HostCallbackOperator:
    """A user-supplied scipy LinearOperator whose matvec is host code: the vector makes a
    round trip per application; every other step of the solver stays on the device."""

    def __init__(self, linear_operator):
        _lib.require_device()
        self.op = linear_operator
        self.dim = linear_operator.shape[0]

    def apply(self, d_in, d_out, stream=None):
        x = download(self.dim, d_in, stream)
        y = numpy.asarray(self.op.dot(x)).reshape(-1)
        upload(y, d_out, stream)


def as_device_operator(operator):
    from openfermion_b200.linalg.linear_qubit_operator import (
        LinearQubitOperator, ParallelLinearQubitOperator, SectorLinearQubitOperator)
    if isinstance(operator, (PlanOperator, SectorOperator, CsrOperator, HostCallbackOperator)):
        return operator
    if isinstance(operator, SectorLinearQubitOperator):
        return SectorOperator(operator.apply_plan, operator.sector)
    if isinstance(operator, LinearQubitOperator):
        return PlanOperator(operator.plan)
    if isinstance(operator, ParallelLinearQubitOperator):
        return PlanOperator(operator._fused.plan)
    if scipy.sparse.issparse(operator):
        return CsrOperator(operator)
    if isinstance(operator, scipy.sparse.linalg.LinearOperator):
        return HostCallbackOperator(operator)
    raise ValueError('operator is not a LinearOperator or sparse matrix: {}.'.format(type(operator)))


def apply_any(operator, state):
    """operator * state for host arrays through the device (used by `expectation`)."""
    dev_op = as_device_operator(operator)
    arr = numpy.asarray(state)
    if arr.size != dev_op.dim:
        raise ValueError('Dimension mismatch')
    vin = _lib.to_device(numpy.ascontiguousarray(arr.reshape(-1), dtype=numpy.complex128))
    vout = _lib.DeviceBuffer(dev_op.dim * C128)
    dev_op.apply(vin.ptr, vout.ptr)
    return download(dev_op.dim, vout.ptr).reshape(arr.shape)


# ----------------------------------------------------------------------------
# Lanczos
# ----------------------------------------------------------------------------
class LocalComm:
    """Reduction hook of the Krylov loops: identity on one GPU; the sharded path
    (openfermion_b200.distributed) substitutes an all-reduce over the ranks."""
    rank = 0
    world = 1

    def allsum(self, value):
        return value


_LOCAL = LocalComm()


def gdot(n, x, y, comm=_LOCAL):
    """conj(x) . y over all shards."""
    return comm.allsum(dotc(n, x, y))


def gnorm(n, x, comm=_LOCAL):
    return float(numpy.sqrt(comm.allsum(dotc(n, x, x)).real))


def _project_out(n, basis_ptrs, w, comm=_LOCAL):
    for u in basis_ptrs:
        axpy(n, -gdot(n, u, w, comm), u, w)


def _lanczos_pass(dev_op, start, locked, krylov_dim, tol, bufs, ritz=None, comm=_LOCAL):
    """One Lanczos sweep from `start` (device pointer, unit norm).

    ritz is None  -> builds T, returns (alphas, betas, theta, s, residual_estimate)
    ritz is given -> replays the same recurrence accumulating y = sum_k ritz[k] v_k into
                     bufs['y'] (the sequence of alpha/beta is recomputed identically).
    """
    n = dev_op.dim
    v_prev, v, w = bufs['a'], bufs['b'], bufs['c']
    copy(n, start, v)
    alphas, betas = [], []
    beta_prev = 0.0
    have_prev = False
    theta, s, res = None, None, None
    steps = krylov_dim if ritz is None else len(ritz)
    if ritz is not None:
        scal(n, 0.0, bufs['y'])
    for it in range(steps):
        if ritz is not None:
            axpy(n, ritz[it], v, bufs['y'])
            if it == steps - 1:
                break
        dev_op.apply(v, w)
        if locked:
            _project_out(n, locked, w, comm)
        alpha = gdot(n, v, w, comm).real
        lanczos_update(n, alpha, beta_prev, v, v_prev if have_prev else None, w)
        beta = gnorm(n, w, comm)
        alphas.append(alpha)
        if ritz is None:
            check = (it < 32) or (it % 4 == 3) or it == steps - 1
            if check:
                d = numpy.array(alphas)
                e = numpy.array(betas)
                vals, vecs = scipy.linalg.eigh_tridiagonal(d, e, select='i', select_range=(0, 0)) \
                    if len(d) > 1 else (d.copy(), numpy.ones((1, 1)))
                theta, s = float(vals[0]), vecs[:, 0]
                res = abs(beta * s[-1])
                scale = max(1.0, abs(theta))
                if res <= tol * scale or beta <= 1e-14 * scale:
                    break
        if it == steps - 1:
            break
        betas.append(beta)
        scal(n, 1.0 / beta, w)
        v_prev, v, w = v, w, v_prev
        have_prev = True
        beta_prev = beta
    return alphas, betas, theta, s, res


def _stored_basis_capacity(dev_op, krylov_dim, comm):
    """Number of Lanczos vectors that can stay in HBM (70 % of the free memory), or 0 to keep
    the two-pass scheme (also with OFB_LANCZOS_STORE=0).  One GPU only: the ranks of a sharded
    solve would have to agree on the number."""
    import os
    if os.environ.get('OFB_LANCZOS_STORE', '1') != '1' or comm.world != 1:
        return 0
    device = getattr(getattr(dev_op, 'plan', None), 'device', 0)
    free = ctypes.c_size_t()
    _lib.call('ofb_device_info', c_int(device), None, None, ctypes.byref(free), None, None)
    fit = int(0.7 * free.value) // (dev_op.dim * C128)
    return min(krylov_dim, fit) if fit >= 16 else 0


def _lanczos_pass_stored(dev_op, start, locked, steps, tol, bufs, basis, comm=_LOCAL):
    """One Lanczos sweep that keeps v_0 .. v_{m-1} as the columns of `basis` (steps columns) and
    forms the Ritz vector y = V s with one pass over the stored vectors instead of replaying the
    recurrence: m matrix-vector products instead of 2 m - 1.  Same recurrence and stopping rule
    as _lanczos_pass."""
    n = dev_op.dim

    def col(k):
        return basis + k * n * C128
    copy(n, start, col(0))
    alphas, betas = [], []
    beta_prev = 0.0
    theta, s = None, None
    for it in range(steps):
        v = col(it)
        w = col(it + 1) if it + 1 < steps else bufs['c']
        dev_op.apply(v, w)
        if locked:
            _project_out(n, locked, w, comm)
        alpha = gdot(n, v, w, comm).real
        lanczos_update(n, alpha, beta_prev, v, col(it - 1) if it else None, w)
        beta = gnorm(n, w, comm)
        alphas.append(alpha)
        if (it < 32) or (it % 4 == 3) or it == steps - 1:
            d, e = numpy.array(alphas), numpy.array(betas)
            vals, vecs = scipy.linalg.eigh_tridiagonal(d, e, select='i', select_range=(0, 0)) \
                if len(d) > 1 else (d.copy(), numpy.ones((1, 1)))
            theta, s = float(vals[0]), vecs[:, 0]
            scale = max(1.0, abs(theta))
            if abs(beta * s[-1]) <= tol * scale or beta <= 1e-14 * scale:
                break
        if it == steps - 1:
            break
        betas.append(beta)
        scal(n, 1.0 / beta, w)
        beta_prev = beta
    _lib.call('ofb_memset', c_void_p(bufs['y']), c_int(0), ctypes.c_size_t(n * C128), c_void_p(None))
    multi_axpy(n, numpy.asarray(s, dtype=numpy.complex128), basis, n, 0.0, bufs['y'])
    return alphas, betas, theta, s


def _lanczos_lowest_native(dev_op, initial_guess, locked, tol, krylov_dim, max_restarts, seed, stats):
    """ofb_lanczos_lowest (include/ofb200.h): restarted Lanczos on unnormalised vectors with all
    scalars on the device, one fused update pass per iteration, alpha from the apply kernel."""
    n = dev_op.dim
    start = None
    if initial_guess is not None:
        guess = numpy.asarray(initial_guess).reshape(-1)
        if guess.size != n:
            raise ValueError('initial guess has the wrong dimension')
        start = _lib.DeviceBuffer(n * C128)
        upload(guess, start.ptr)
    out = _lib.DeviceBuffer(n * C128)
    locked = list(locked)
    locked_arr = (c_void_p * max(len(locked), 1))(*locked)
    energy, residual = c_double(), c_double()
    matvecs, restarts = c_int64(), c_int()
    store = 1 if os.environ.get('OFB_LANCZOS_STORE', '1') == '1' else 0
    op = dev_op.linop()
    try:
        _lib.call('ofb_lanczos_lowest', op, c_void_p(start.ptr if start else None), c_int(len(locked)),
                  locked_arr, c_double(tol), c_int(krylov_dim or 0), c_int(max_restarts),
                  ctypes.c_uint64(seed), c_int(store), ctypes.byref(energy), ctypes.byref(residual),
                  c_void_p(out.ptr), ctypes.byref(matvecs), ctypes.byref(restarts), c_void_p(None))
    finally:
        _lib.call('ofb_linop_destroy', op)
        if start is not None:
            start.free()
    if stats is not None:
        stats['matvecs'] = stats.get('matvecs', 0) + matvecs.value
        stats['restarts'] = restarts.value
    return energy.value, out, residual.value


def lanczos_lowest(dev_op, initial_guess=None, locked=(), tol=1e-10, krylov_dim=None,
                   max_restarts=40, seed=1234, comm=_LOCAL, stats=None):
    """Lowest eigenpair of a Hermitian device operator, orthogonal to `locked` vectors.

    Returns (eigenvalue, device buffer holding the unit eigenvector, residual norm).
    """
    n = dev_op.dim
    if comm.world == 1 and hasattr(dev_op, 'linop') and os.environ.get('OFB_LANCZOS_HOST') != '1':
        # one device: the whole restarted solver runs behind the C ABI (csrc/krylov.cu)
        return _lanczos_lowest_native(dev_op, initial_guess, locked, tol, krylov_dim, max_restarts,
                                      seed, stats)
    if krylov_dim is None:
        krylov_dim = 400
    krylov_dim = max(2, min(krylov_dim, n * comm.world - len(locked)))
    owners = {k: _lib.DeviceBuffer(n * C128) for k in 'abcy'}
    bufs = {k: b.ptr for k, b in owners.items()}
    start = _lib.DeviceBuffer(n * C128)
    if initial_guess is not None:
        guess = numpy.asarray(initial_guess).reshape(-1)
        if guess.size != n:
            raise ValueError('initial guess has the wrong dimension')
        upload(guess, start.ptr)
    else:
        fill_random(n, seed + 104729 * comm.rank, False, start.ptr)
    hy = _lib.DeviceBuffer(n * C128)
    stored = _stored_basis_capacity(dev_op, krylov_dim, comm)
    basis = _lib.DeviceBuffer(stored * n * C128) if stored else None
    theta = 0.0
    residual = numpy.inf
    for restart in range(max_restarts):
        if locked:
            _project_out(n, locked, start.ptr, comm)
        norm = gnorm(n, start.ptr, comm)
        if norm == 0.0:
            fill_random(n, seed + 7919 * (restart + 1) + 104729 * comm.rank, False, start.ptr)
            continue
        scal(n, 1.0 / norm, start.ptr)
        if basis is not None:
            alphas, betas, theta, s = _lanczos_pass_stored(dev_op, start.ptr, locked, stored, tol, bufs,
                                                           basis.ptr, comm=comm)
            applied = len(alphas)
        else:
            alphas, betas, theta, s, _ = _lanczos_pass(dev_op, start.ptr, locked, krylov_dim, tol, bufs,
                                                       comm=comm)
            _lanczos_pass(dev_op, start.ptr, locked, krylov_dim, tol, bufs, ritz=s, comm=comm)
            applied = 2 * len(alphas) - 1
        if stats is not None:
            stats['matvecs'] = stats.get('matvecs', 0) + applied + 1
            stats['restarts'] = restart + 1
        y = bufs['y']
        if locked:
            _project_out(n, locked, y, comm)
        scal(n, 1.0 / gnorm(n, y, comm), y)
        dev_op.apply(y, hy.ptr)
        theta = gdot(n, y, hy.ptr, comm).real
        axpy(n, -theta, y, hy.ptr)
        residual = gnorm(n, hy.ptr, comm)
        copy(n, y, start.ptr)
        if residual <= 10.0 * tol * max(1.0, abs(theta)):
            break
    for owner in owners.values():
        owner.free()
    hy.free()
    if basis is not None:
        basis.free()
    return theta, start, residual


def lowest_eigenpairs(operator, k, initial_guess=None, tol=1e-10):
    """k lowest eigenpairs by Lanczos with locking (k = 1 for get_ground_state, 2 for get_gap).
    Returns (float64[k], complex128[dim, k])."""
    dev_op = as_device_operator(operator)
    n = dev_op.dim
    if initial_guess is not None:
        initial_guess = numpy.asarray(initial_guess)
    values, holders = [], []
    for j in range(min(k, n)):
        guess = None
        if initial_guess is not None:
            guess = initial_guess if initial_guess.ndim == 1 else initial_guess[:, min(j, initial_guess.shape[1] - 1)]
            if j > 0 and initial_guess.ndim == 1:
                guess = None
        theta, vec, _ = lanczos_lowest(dev_op, guess, [h.ptr for h in holders], tol, seed=1234 + j)
        values.append(theta)
        holders.append(vec)
    vectors = numpy.stack([download(n, h.ptr) for h in holders], axis=1)
    order = numpy.argsort(values)
    return numpy.array(values)[order], vectors[:, order]
