"""Restarted Lanczos on a state sharded over GPUs (BASELINE.json config 5), driven from Python
with the per-iteration primitives of csrc/krylov.cu (include/ofb200.h, ofb_kws_*).

Per iteration and rank:  H u  (K1 launches of the sharded operator; the launch that completes the
output shard also leaves the local part of <u, H u>)  ->  all-reduce of 2 doubles  ->  ofb_kws_alpha
->  ofb_kws_step (ONE fused pass: w/beta - (alpha/beta) u - (beta/beta') u', local |w|^2)  ->
all-reduce of 1 double  ->  ofb_kws_beta.  Nothing in an iteration synchronises the host: the
scalars live in the workspace's device slots, NCCL all-reduces them in stream order, and the
host only fetches the alpha/beta history when it checks convergence.

Cross-rank ordering in the 'ce' / 'p2p' exchange modes: a rank reads its peers' copies of u_k.
The all-reduce that follows the step kernel completes on a rank's stream only after every
rank's step kernel has finished, so the peers' u_{k+1} is complete before the next apply reads
it; the Ritz-vector pass (no reductions) inserts a one-double all-reduce for the same purpose.
A buffer is overwritten two iterations after it was last read remotely, again behind an
all-reduce that the reader joins only after its reads.  All buffers an apply can read are
registered with the peers once and unregistered before they are freed (`close`).

The reference has no counterpart (get_ground_state, linalg/sparse_tools.py:566-587, is ARPACK on
one host); results are compared with the one-GPU solver and the reference's energies in
tests/dist_gpu_check.py.
"""
import ctypes

import numpy

from openfermion_b200 import _lib
from openfermion_b200._lib import c_double, c_int, c_void_p

C128 = 16
SLOT_DOT, SLOT_NRM2, SLOT_SYNC = 0, 2, 12


class ShardedLanczos:
    def __init__(self, operator, dist, group=None, max_steps=400, vectors=5):
        """vectors: 5 = full solver (three recurrence vectors, Ritz vector, restart vector);
        3 = recurrence only (`iterate`, the per-iteration timing of bench.py)."""
        import torch
        self.op = operator
        self.dist = dist
        self.group = group
        self.torch = torch
        self.n = operator.dim
        self.max_steps = int(max_steps)
        self.device = operator._device_index()
        self.ws = c_void_p()
        _lib.call('ofb_kws_create', ctypes.c_int64(self.n), c_int(self.max_steps), c_int(self.device),
                  ctypes.byref(self.ws))
        ptr = c_void_p()
        _lib.call('ofb_kws_scalars', self.ws, ctypes.byref(ptr))

        class _Iface:
            pass
        holder = _Iface()
        holder.__cuda_array_interface__ = {'shape': (16,), 'typestr': '<f8', 'data': (ptr.value, False),
                                           'version': 2}
        self.scal = torch.as_tensor(holder, device='cuda:%d' % self.device)
        self.bufs = [operator.new_vector() for _ in range(vectors)]
        self.exchanging = operator.mode in ('ce', 'p2p') and bool(operator.layout.remote_patterns)
        if self.exchanging:
            operator.register_peer_vectors(self.bufs)
        self._closed = False

    # -- plumbing -----------------------------------------------------------------------------
    def _stream(self):
        return c_void_p(self.op._stream())

    def _allreduce(self, lo, hi):
        self.dist.all_reduce(self.scal[lo:hi], group=self.group)

    def _allsum_host(self, value):
        t = self.torch.tensor([value.real, value.imag], dtype=self.torch.float64,
                              device='cuda:%d' % self.device)
        self.dist.all_reduce(t, group=self.group)
        out = t.cpu()
        return complex(float(out[0]), float(out[1]))

    def _gdot(self, x, y):
        from openfermion_b200.linalg import krylov
        return self._allsum_host(krylov.dotc(self.n, x.ptr, y.ptr, self.op._stream()))

    def _gnorm(self, x):
        return float(numpy.sqrt(self._gdot(x, x).real))

    def _copy(self, src, dst):
        _lib.call('ofb_memcpy_d2d', c_void_p(dst.ptr), c_void_p(src.ptr), ctypes.c_size_t(self.n * C128),
                  self._stream())

    def _scale(self, x, factor):
        from openfermion_b200.linalg import krylov
        krylov.scal(self.n, factor, x.ptr, self.op._stream())

    def _fetch(self, count):
        alphas = numpy.zeros(max(count, 1))
        betas = numpy.zeros(count + 1)
        flag = c_int()
        _lib.call('ofb_kws_fetch', self.ws, c_int(count), _lib.ptr(alphas), _lib.ptr(betas),
                  ctypes.byref(flag), self._stream())
        return alphas[:count], betas, flag.value

    @staticmethod
    def _ritz(alphas, betas, m):
        value = c_double()
        vector = numpy.zeros(m)
        off = numpy.ascontiguousarray(betas[1:m]) if m > 1 else numpy.zeros(1)
        _lib.call('ofb_tridiag_lowest', c_int(m), _lib.ptr(numpy.ascontiguousarray(alphas[:m])),
                  _lib.ptr(off), ctypes.byref(value), _lib.ptr(vector))
        return value.value, vector

    # -- one iteration --------------------------------------------------------------------------------
    def _iteration(self, k, u_prev, u, w, y=None, replay=False, ritz_k=0.0):
        """u_{k+1} into w (in place of H u_k).  First pass: fused dot and norm, two all-reduces."""
        s = self._stream()
        if replay:
            if self.exchanging:  # orders the peers' writes of u_k before our reads (no reduction here)
                self._allreduce(SLOT_SYNC, SLOT_SYNC + 1)
            self.op.matvec(u, w)
        else:
            self.op.matvec(u, w, dot=(self.ws, u))
            self._allreduce(SLOT_DOT, SLOT_DOT + 2)
        _lib.call('ofb_kws_alpha', self.ws, c_int(k), c_int(1 if replay else 0), c_double(ritz_k), s)
        _lib.call('ofb_kws_step', self.ws, c_void_p(w.ptr), c_void_p(u.ptr),
                  c_void_p(u_prev.ptr if (k and u_prev is not None) else None),
                  c_void_p(y.ptr if y is not None else None), c_int(1 if replay else 0), s)
        if not replay:
            self._allreduce(SLOT_NRM2, SLOT_NRM2 + 1)
            _lib.call('ofb_kws_beta', self.ws, c_int(k), s)

    def _begin(self, start_into):
        """Common start of a sweep: the (unit-norm) start vector sits in `start_into`."""
        _lib.call('ofb_kws_reset', self.ws, c_double(1.0), self._stream())
        if self.exchanging:  # every rank's start shard is complete before anyone reads a peer's
            self.dist.barrier(group=self.group)

    def prepare(self, seed=1234):
        """Seeded global random start vector, normalised; resets the recurrence."""
        a, b, c = self.bufs[:3]
        self.op.fill_global_random(b, seed)
        self._scale(b, 1.0 / self._gnorm(b))
        self._begin(b)
        self._state = [0, a, b, c]

    def run(self, steps):
        """`steps` more first-pass iterations, queued WITHOUT any host synchronisation (bench.py
        times exactly this between two CUDA events)."""
        k, up, u, w = self._state
        for _ in range(steps):
            self._iteration(k, up, u, w)
            up, u, w = u, w, up
            k += 1
        self._state = [k, up, u, w]

    def history(self):
        """(alphas, betas, lowest Ritz value) of the iterations run so far (synchronises)."""
        count = self._state[0]
        alphas, betas, flag = self._fetch(count)
        m = min(count, flag) if flag else count
        theta, _ = self._ritz(alphas, betas, m)
        return alphas, betas, theta

    # -- the solver ---------------------------------------------------------------------------------------
    def solve(self, tol=1e-10, max_restarts=40, seed=1234, stats=None):
        """Lowest eigenpair.  Returns (energy, buffer with this rank's shard of the unit
        eigenvector -- owned by the caller --, residual norm)."""
        from openfermion_b200.linalg import krylov
        if len(self.bufs) < 5:
            raise RuntimeError('solver needs 5 vectors')
        a, b, c, y, start = self.bufs
        steps_max = max(2, min(self.max_steps, self.n * self.op.world))
        self.op.fill_global_random(start, seed)
        theta, residual, matvecs = 0.0, numpy.inf, 0
        restart = 0
        for restart in range(max_restarts):
            norm = self._gnorm(start)
            if norm == 0.0:
                self.op.fill_global_random(start, seed + 7919 * (restart + 1))
                continue
            self._scale(start, 1.0 / norm)
            # first pass
            self._copy(start, b)
            self._begin(b)
            up, u, w = a, b, c
            m, ritz = 0, None
            for k in range(steps_max):
                self._iteration(k, up, u, w)
                up, u, w = u, w, up
                m = k + 1
                if k < 8 or (k & 3) == 3 or k == steps_max - 1:
                    alphas, betas, flag = self._fetch(m)
                    if flag:
                        m = min(m, flag)
                    theta, ritz = self._ritz(alphas, betas, m)
                    scale = max(1.0, abs(theta))
                    if flag or abs(betas[m] * ritz[m - 1]) <= tol * scale or betas[m] <= 1e-14 * scale:
                        break
            matvecs += m
            # second pass: the same recurrence from its recorded coefficients, y = sum_k s_k v_k
            _lib.call('ofb_memset', c_void_p(y.ptr), c_int(0), ctypes.c_size_t(self.n * C128), self._stream())
            self._copy(start, b)
            if self.exchanging:
                self.dist.barrier(group=self.group)
            up, u, w = a, b, c
            for k in range(m):
                if k == m - 1:  # the last vector only enters y
                    s = self._stream()
                    _lib.call('ofb_kws_alpha', self.ws, c_int(k), c_int(1), c_double(float(ritz[k])), s)
                    _lib.call('ofb_kws_step', self.ws, c_void_p(w.ptr), c_void_p(u.ptr), c_void_p(None),
                              c_void_p(y.ptr), c_int(1), s)
                    break
                self._iteration(k, up, u, w, y=y, replay=True, ritz_k=float(ritz[k]))
                up, u, w = u, w, up
                matvecs += 1
            self._scale(y, 1.0 / self._gnorm(y))
            # true residual; the Rayleigh quotient is the returned energy
            if self.exchanging:
                self.dist.barrier(group=self.group)
            self.op.matvec(y, a)
            matvecs += 1
            theta = self._gdot(y, a).real
            krylov.axpy(self.n, -theta, y.ptr, a.ptr, self.op._stream())
            residual = self._gnorm(a)
            self._copy(y, start)
            if residual <= 10.0 * tol * max(1.0, abs(theta)):
                break
        if stats is not None:
            stats['matvecs'] = stats.get('matvecs', 0) + matvecs
            stats['restarts'] = restart + 1
        result = self.op.new_vector()
        self._copy(start, result)
        self.torch.cuda.synchronize()
        return theta, result, residual

    def close(self):
        if self._closed:
            return
        self._closed = True
        self.torch.cuda.synchronize()
        if self.exchanging:
            self.op.unregister_peer_vectors(self.bufs)
        for buf in self.bufs:
            buf.free()
        self.bufs = []
        _lib.call('ofb_kws_destroy', self.ws)
