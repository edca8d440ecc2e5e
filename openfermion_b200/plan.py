"""Python owner of an `ofb_plan` (compiled term tables on one GPU)."""
import ctypes
import os

import numpy
import scipy.sparse

from openfermion_b200 import _lib
from openfermion_b200._lib import c_void_p, c_int, c_int64


class PauliPlan:
    """Compiled Pauli sum: apply / expectation / diagonal / CSC through the C ABI."""

    def __init__(self, n_qubits, x_mask, z_mask, coeff, device=0, hints=True):
        _lib.require_device()
        self.n_qubits = int(n_qubits)
        self.dim = 1 << self.n_qubits
        x = numpy.ascontiguousarray(x_mask, dtype=numpy.uint64)
        z = numpy.ascontiguousarray(z_mask, dtype=numpy.uint64)
        c = numpy.ascontiguousarray(coeff, dtype=numpy.complex128)
        if not (len(x) == len(z) == len(c)):
            raise ValueError('term table arrays differ in length')
        handle = c_void_p()
        _lib.call('ofb_plan_create', c_int(self.n_qubits), c_int64(len(x)), _lib.ptr(x), _lib.ptr(z),
                  _lib.ptr(c), c_int(device), ctypes.byref(handle))
        self.handle = handle
        self.device = device
        self.projected = False
        self._read_info()
        self.support_fraction = 1.0
        self._k1_handle = None  # separate plan of the reduced term table (support hints)
        # Support hints (openfermion_b200/support.py) are on by default: K1 / K2 skip the gathers
        # on which a group's coefficient is structurally zero and run on the orbit-reduced term
        # table.  OFB_SUPPORT_HINTS=0 turns them off, =hints keeps the full table.
        mode = os.environ.get('OFB_SUPPORT_HINTS', '1')
        if mode in ('1', 'hints') and self.n_qubits <= 32 and len(c) and hints:
            self.set_support_hints(x, z, c, reduce=(mode == '1'))

    @classmethod
    def projected_rows(cls, n_qubits, occ_mask, occ_val, x_mask, z_mask, coeff, device=0):
        """Fermion-route plan (jordan_wigner_sparse); only the CSC calls are valid."""
        _lib.require_device()
        self = cls.__new__(cls)
        self.n_qubits = int(n_qubits)
        self.dim = 1 << self.n_qubits
        arrays = [numpy.ascontiguousarray(a, dtype=numpy.uint64)
                  for a in (occ_mask, occ_val, x_mask, z_mask)]
        c = numpy.ascontiguousarray(coeff, dtype=numpy.complex128)
        handle = c_void_p()
        _lib.call('ofb_plan_create_projected', c_int(self.n_qubits), c_int64(len(c)),
                  *[_lib.ptr(a) for a in arrays], _lib.ptr(c), c_int(device), ctypes.byref(handle))
        self.handle = handle
        self.device = device
        self.projected = True
        self._read_info()
        return self

    def _read_info(self):
        n = c_int()
        t, g, d = c_int64(), c_int64(), c_int64()
        dev = c_int()
        _lib.call('ofb_plan_info', self.handle, ctypes.byref(n), ctypes.byref(t), ctypes.byref(g),
                  ctypes.byref(d), ctypes.byref(dev))
        self.n_terms, self.n_groups, self.n_diag_terms = t.value, g.value, d.value

    def set_support_hints(self, x_mask, z_mask, coeff, reduce=True):
        """Support hints (openfermion_b200/support.py): tell the kernels where each group's
        coefficient is structurally zero, so that those gathers are skipped; with reduce=True
        K1/K2 also run on the orbit-reduced term table (one term per symmetry orbit), held in
        a second library plan -- diagonal and CSC keep the full table."""
        from openfermion_b200 import support
        group_x, mask1, mask2, flags = support.group_conditions(x_mask, z_mask, coeff)
        if len(group_x) != self.n_groups:
            raise ValueError('support hints were derived for a different term table')
        m1 = numpy.ascontiguousarray(mask1, dtype=numpy.uint32)
        m2 = numpy.ascontiguousarray(mask2, dtype=numpy.uint32)
        flags = numpy.ascontiguousarray(flags, dtype=numpy.uint8)
        target = self.handle
        if reduce and numpy.any(flags):
            xr, zr, cr = support.reduce_terms(x_mask, z_mask, coeff, group_x, mask1, mask2, flags)
            xr, zr, cr = (numpy.ascontiguousarray(a) for a in (xr, zr, cr))
            self._drop_k1()
            handle = c_void_p()
            _lib.call('ofb_plan_create', c_int(self.n_qubits), c_int64(len(xr)), _lib.ptr(xr), _lib.ptr(zr),
                      _lib.ptr(cr), c_int(self.device), ctypes.byref(handle))
            self._k1_handle = target = handle
            self.n_reduced_terms = len(xr)
        _lib.call('ofb_plan_set_support', target, c_int64(self.n_groups), _lib.ptr(m1), _lib.ptr(m2),
                  _lib.ptr(flags))
        self.support_fraction = support.needed_fraction(flags)

    def _drop_k1(self):
        if getattr(self, '_k1_handle', None):
            _lib.call('ofb_plan_destroy', self._k1_handle)
            self._k1_handle = None

    @property
    def _apply_handle(self):
        """The plan K1 / K2 run on: the reduced one when support hints are active."""
        return getattr(self, '_k1_handle', None) or self.handle

    def group(self, g):
        x = ctypes.c_uint64()
        count = c_int64()
        _lib.call('ofb_plan_group', self.handle, c_int64(g), ctypes.byref(x), ctypes.byref(count))
        return x.value, count.value

    def group_x_masks(self):
        return numpy.array([self.group(g)[0] for g in range(self.n_groups)], dtype=numpy.uint64)

    # ---- host-buffer calls (the ones the drop-in API makes) -------------------
    # OFB_PIN_THRESHOLD=<bytes>: pin host arrays at least this large in place (cudaHostRegister)
    # around the call.  Off by default: measured on B200 at 28 qubits (4.3 GB each way), one call
    # takes 2.33 s with plain pageable copies and 2.89 s with in-place pinning -- registering and
    # unregistering 8.6 GB costs more than the faster copies save.
    PIN_THRESHOLD = int(os.environ.get('OFB_PIN_THRESHOLD', '0'))

    def apply_host(self, columns):
        """columns: complex128, Fortran-contiguous (dim, k) or (dim,) -> same shape."""
        cols = numpy.asarray(columns, dtype=numpy.complex128)
        ncols = 1 if cols.ndim == 1 else cols.shape[1]
        cols = numpy.asfortranarray(cols) if cols.ndim == 2 else numpy.ascontiguousarray(cols)
        out = numpy.empty_like(cols, order='F' if cols.ndim == 2 else 'C')
        pinned = []
        if 0 < self.PIN_THRESHOLD <= cols.nbytes:
            for array in (cols, out):
                try:
                    _lib.call('ofb_host_register', _lib.ptr(array), ctypes.c_size_t(array.nbytes))
                    pinned.append(array)
                except _lib.OfbError:  # locked-memory limit, already registered ...: plain copies
                    pass
        try:
            _lib.call('ofb_apply_host', self._apply_handle, _lib.ptr(cols), _lib.ptr(out), c_int(ncols))
        finally:
            for array in pinned:
                _lib.call('ofb_host_unregister', _lib.ptr(array))
        return out

    def expectation_host(self, state):
        state = numpy.ascontiguousarray(state, dtype=numpy.complex128).reshape(-1)
        out = (ctypes.c_double * 2)()
        _lib.call('ofb_expectation_host', self._apply_handle, _lib.ptr(state), out)
        return complex(out[0], out[1])

    def diagonal_host(self):
        out = numpy.empty(self.dim, dtype=numpy.float64)
        _lib.call('ofb_diagonal_host', self.handle, _lib.ptr(out))
        return out

    # CSC outputs at least this large leave the device through _lib.parallel_d2h
    PARALLEL_D2H_BYTES = int(os.environ.get('OFB_PARALLEL_D2H_BYTES', str(1 << 26)))
    # term-order sums are bit-identical to the reference up to this many triplets per column
    # (libstdc++'s std::sort is a stable insertion sort up to 16 elements)
    STABLE_SORT_LIMIT = 16
    # auto mode replays scipy's sort order while n_terms * 2^n stays below this
    EXACT_MODE_BUDGET = 1 << 31

    def csc_mode(self, mode=None):
        """0 = term-order sums, 1 = scipy sort-order emulation (include/ofb200.h,
        ofb_plan_set_csc_mode).  `mode` / $OFB_CSC_MODE: 'fast', 'exact' or 'auto' (default):
        exact whenever it can differ from the fast mode and the replay is affordable."""
        import os
        mode = mode or os.environ.get('OFB_CSC_MODE', 'auto')
        if mode == 'fast':
            return 0
        if mode == 'exact':
            return 1
        if mode != 'auto':
            raise ValueError('csc mode must be fast, exact or auto')
        needs = self.n_terms > self.STABLE_SORT_LIMIT
        return 1 if needs and self.n_terms * self.dim <= self.EXACT_MODE_BUDGET else 0

    def csc_host(self, mode=None):
        """Materialises the operator as scipy CSC with scipy's own index-width rule."""
        _lib.call('ofb_plan_set_csc_mode', self.handle, c_int(self.csc_mode(mode)))
        nnz = c_int64()
        bits = c_int()
        _lib.call('ofb_csc_count', self.handle, ctypes.byref(nnz), ctypes.byref(bits), c_void_p(None))
        idx_dtype = numpy.int32 if bits.value == 32 else numpy.int64
        indptr = numpy.empty(self.dim + 1, dtype=idx_dtype)
        indices = numpy.empty(nnz.value, dtype=idx_dtype)
        data = numpy.empty(nnz.value, dtype=numpy.complex128)
        if indices.nbytes + data.nbytes >= self.PARALLEL_D2H_BYTES:
            # large output (config 2: 15 GB): fill on the device, then staged copies from several
            # worker threads -- a single pageable copy into fresh memory runs at ~5 GB/s because
            # the first touch of every destination page happens inside it
            d_indptr = _lib.DeviceBuffer(indptr.nbytes)
            d_indices = _lib.DeviceBuffer(indices.nbytes)
            d_data = _lib.DeviceBuffer(data.nbytes)
            try:
                _lib.call('ofb_csc_fill', self.handle, c_void_p(d_indptr.ptr), c_void_p(d_indices.ptr),
                          c_void_p(d_data.ptr), bits, c_void_p(None))
                _lib.call('ofb_device_sync')
                _lib.parallel_d2h([(indptr, d_indptr.ptr), (indices, d_indices.ptr), (data, d_data.ptr)],
                                  device=self.device)
            finally:
                for buf in (d_indptr, d_indices, d_data):
                    buf.free()
        else:
            _lib.call('ofb_csc_fill_host', self.handle, _lib.ptr(indptr), _lib.ptr(indices),
                      _lib.ptr(data), bits)
        if nnz.value == 0:
            # With no terms at all the reference concatenates only its empty float seed lists
            # and returns a float64 matrix; with terms that cancel it stays complex128
            # (SURVEY.md section 7, quirk (c)).
            dtype = numpy.float64 if self.n_terms == 0 else numpy.complex128
            return scipy.sparse.csc_matrix((self.dim, self.dim), dtype=dtype)
        mat = scipy.sparse.csc_matrix((data, indices, indptr), shape=(self.dim, self.dim))
        mat.has_sorted_indices = True
        return mat

    # ---- device-pointer calls (Krylov loops, bench) ----------------------------
    def apply_device(self, d_in, d_out, ncols=1, ld=None, stream=None):
        _lib.call('ofb_apply', self._apply_handle, c_void_p(d_in), c_void_p(d_out), c_int(ncols),
                  c_int64(self.dim if ld is None else ld), c_void_p(stream))

    def apply_groups_device(self, d_in, d_out, local_bits, index_base, g_begin, g_end, accumulate,
                            stream=None):
        _lib.call('ofb_apply_groups', self._apply_handle, c_void_p(d_in), c_void_p(d_out), c_int(local_bits),
                  ctypes.c_uint64(index_base), c_int64(g_begin), c_int64(g_end), c_int(accumulate),
                  c_void_p(stream))

    def expectation_device(self, d_state, stream=None):
        out = (ctypes.c_double * 2)()
        _lib.call('ofb_expectation', self._apply_handle, c_void_p(d_state), out, c_void_p(stream))
        return complex(out[0], out[1])

    def diagonal_device(self, d_out, stream=None):
        _lib.call('ofb_diagonal', self.handle, c_void_p(d_out), c_void_p(stream))

    def close(self):
        self._drop_k1()
        if getattr(self, 'handle', None):
            _lib.call('ofb_plan_destroy', self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
