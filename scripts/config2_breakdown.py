#!/usr/bin/env python
"""Where the time of config 2 (qubit_operator_sparse at 20 qubits, 744 M non-zeros, 15 GB) goes:
term compilation, plan, count, fill, device -> host copy (threads x pre-touch sweep), scipy wrap."""
import ctypes
import json
import os
import sys
import time

import numpy
import scipy.sparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, compiler, hamiltonians  # noqa: E402
from openfermion_b200.linalg import qubit_operator_sparse  # noqa: E402
from openfermion_b200.plan import PauliPlan  # noqa: E402

c_int, c_int64, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p


def main():
    n = 20
    _lib.require_device()
    t = time.perf_counter()
    op = hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16)
    out = {'operator_s': time.perf_counter() - t}
    t = time.perf_counter()
    x, z, c = compiler.compile_qubit_operator(op, n)
    out['compile_s'] = time.perf_counter() - t
    t = time.perf_counter()
    plan = PauliPlan(n, x, z, c)
    _lib.call('ofb_plan_set_csc_mode', plan.handle, c_int(0))
    out['plan_s'] = time.perf_counter() - t
    nnz, bits = c_int64(), c_int()
    t = time.perf_counter()
    _lib.call('ofb_csc_count', plan.handle, ctypes.byref(nnz), ctypes.byref(bits), vp(None))
    out['count_s'] = time.perf_counter() - t
    dim = 1 << n
    t = time.perf_counter()
    d_indptr = _lib.DeviceBuffer((dim + 1) * 4)
    d_indices = _lib.DeviceBuffer(nnz.value * 4)
    d_data = _lib.DeviceBuffer(nnz.value * 16)
    out['device_alloc_s'] = time.perf_counter() - t
    t = time.perf_counter()
    _lib.call('ofb_csc_fill', plan.handle, vp(d_indptr.ptr), vp(d_indices.ptr), vp(d_data.ptr), bits, vp(None))
    _lib.call('ofb_device_sync')
    out['fill_s'] = time.perf_counter() - t
    out['d2h'] = []
    for threads, pretouch, staged in ((1, False, False), (8, False, False), (8, True, False), (16, True, False),
                                      (2, False, True), (4, False, True), (8, False, True), (16, False, True)):
        if True:
            t = time.perf_counter()
            indptr = numpy.empty(dim + 1, dtype=numpy.int32)
            indices = numpy.empty(nnz.value, dtype=numpy.int32)
            data = numpy.empty(nnz.value, dtype=numpy.complex128)
            t1 = time.perf_counter()
            _lib.parallel_d2h([(indptr, d_indptr.ptr), (indices, d_indices.ptr), (data, d_data.ptr)],
                              threads=threads, pretouch=pretouch, staged=staged)
            t2 = time.perf_counter()
            mat = scipy.sparse.csc_matrix((data, indices, indptr), shape=(dim, dim))
            t3 = time.perf_counter()
            del mat, data, indices, indptr
            t4 = time.perf_counter()
            out['d2h'].append({'threads': threads, 'pretouch': pretouch, 'staged': staged, 'empty_s': t1 - t, 'copy_s': t2 - t1,
                               'GBs': (nnz.value * 20 + dim * 4) / (t2 - t1) / 1e9, 'scipy_wrap_s': t3 - t2,
                               'free_s': t4 - t3})
            print(json.dumps(out['d2h'][-1]), flush=True)
    # host -> device, 4 GB of touched pageable memory: plain copy against the staged one
    src = numpy.ones(1 << 28, dtype=numpy.complex128)
    out['h2d'] = []
    for threads in (0, 2, 4, 8, 16):
        t = time.perf_counter()
        if threads == 0:
            _lib.call('ofb_memcpy_h2d', vp(d_data.ptr), _lib.ptr(src), ctypes.c_size_t(src.nbytes), vp(None))
            _lib.call('ofb_stream_sync', vp(None))
        else:
            _lib.call('ofb_copy_h2d_staged', vp(d_data.ptr), _lib.ptr(src), ctypes.c_size_t(src.nbytes), c_int(0),
                      c_int(threads))
        dt = time.perf_counter() - t
        out['h2d'].append({'threads': threads, 's': dt, 'GBs': src.nbytes / dt / 1e9})
        print(json.dumps(out['h2d'][-1]), flush=True)
    back = numpy.empty_like(src)
    _lib.call('ofb_copy_d2h_staged', _lib.ptr(back), vp(d_data.ptr), ctypes.c_size_t(src.nbytes), c_int(0), c_int(8))
    assert numpy.array_equal(back, src)
    del src, back
    for b in (d_indptr, d_indices, d_data):
        b.free()
    plan.close()
    for k in range(2):
        t = time.perf_counter()
        mat = qubit_operator_sparse(op, n)
        out['qubit_operator_sparse_s_%d' % k] = time.perf_counter() - t
        del mat
    print(json.dumps(out))


if __name__ == '__main__':
    main()
