#!/usr/bin/env python
"""Device-side timing of the CSC builder (count + scan, fill) without the device->host copy."""
import ctypes
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, compiler, hamiltonians  # noqa: E402
from openfermion_b200.plan import PauliPlan  # noqa: E402

c_int, c_int64, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p


def run(name, op, n, mode):
    x, z, c = compiler.compile_qubit_operator(op, n)
    plan = PauliPlan(n, x, z, c)
    _lib.call('ofb_plan_set_csc_mode', plan.handle, c_int(mode))
    nnz, bits = c_int64(), c_int()
    _lib.call('ofb_device_sync')
    t0 = time.perf_counter()
    _lib.call('ofb_csc_count', plan.handle, ctypes.byref(nnz), ctypes.byref(bits), vp(None))
    _lib.call('ofb_device_sync')
    t1 = time.perf_counter()
    ib = bits.value // 8
    dim = 1 << n
    d_indptr = _lib.DeviceBuffer((dim + 1) * ib)
    d_indices = _lib.DeviceBuffer(max(nnz.value, 1) * ib)
    d_data = _lib.DeviceBuffer(max(nnz.value, 1) * 16)
    _lib.call('ofb_device_sync')
    t2 = time.perf_counter()
    _lib.call('ofb_csc_fill', plan.handle, vp(d_indptr.ptr), vp(d_indices.ptr), vp(d_data.ptr), bits, vp(None))
    _lib.call('ofb_device_sync')
    t3 = time.perf_counter()
    out_bytes = nnz.value * (16 + ib) + (dim + 1) * ib
    print(json.dumps({'case': name, 'n_qubits': n, 'terms': len(c), 'groups': plan.n_groups,
                      'mode': 'scipy-order' if mode else 'term-order', 'nnz': nnz.value,
                      'count_scan_s': t1 - t0, 'fill_s': t3 - t2, 'output_GB': out_bytes / 1e9,
                      'fill_GBs': out_bytes / (t3 - t2) / 1e9,
                      'term_column_evals_per_s': 2 * len(c) * dim / ((t1 - t0) + (t3 - t2))}), flush=True)
    plan.close()


def sweep(which):
    """slice counts of the tiled passes (OFB_CSC_COUNT_SLICES / OFB_CSC_FILL_SLICES)"""
    cases = {'mol20': ('molecular dyadic', lambda: hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16), 20),
             'hub24': ('hubbard 3x4', lambda: hamiltonians.hubbard_jw(3, 4), 24),
             'mol16': ('molecular', lambda: hamiltonians.molecular_jw(8, seed=1234), 16)}
    name, make, n = cases[which]
    op = make()
    for s in [int(t) for t in os.environ.get('CSC_SWEEP_LIST', '0,1,2,4,8,16,32,64,128,256').split(',')]:
        for key in ('OFB_CSC_COUNT_SLICES', 'OFB_CSC_FILL_SLICES'):
            if s:
                os.environ[key] = str(s)
            else:
                os.environ.pop(key, None)
        for _ in range(2):
            run('%s slices=%s' % (name, s or 'auto'), op, n, 0)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--sweep':
        sweep(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == '--one':  # a single build, for ncu
        sweep.__globals__['run'](*{'mol20': ('molecular dyadic', hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16), 20, 0),
                                   'hub24': ('hubbard 3x4', hamiltonians.hubbard_jw(3, 4), 24, 0)}[sys.argv[2]])
        sys.exit(0)
    run('molecular dyadic', hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16), 20, 0)
    run('molecular', hamiltonians.molecular_jw(7, seed=1234), 14, 1)
    run('molecular', hamiltonians.molecular_jw(7, seed=1234), 14, 0)
    run('molecular', hamiltonians.molecular_jw(8, seed=1234), 16, 1)
    run('hubbard 2x4', hamiltonians.hubbard_jw(2, 4), 16, 1)
    run('hubbard 2x5', hamiltonians.hubbard_jw(2, 5), 20, 1)
    run('hubbard 3x4', hamiltonians.hubbard_jw(3, 4), 24, 0)
