#!/usr/bin/env python
"""K1b timing: ms per column of the batched H|psi> (ofb_apply with ncols = k) against the
single-column kernel, device-resident columns, CUDA events.   python scripts/multi_bench.py mol26 mol28"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import LinearQubitOperator, krylov  # noqa: E402

builders = {'mol24': lambda: hamiltonians.molecular_jw(12), 'mol26': lambda: hamiltonians.molecular_jw(13),
            'mol28': lambda: hamiltonians.molecular_jw(14), 'hub24': lambda: hamiltonians.hubbard_jw(3, 4),
            'hub28': lambda: hamiltonians.hubbard_jw(2, 7)}
vp = ctypes.c_void_p
ev = [vp(), vp()]
for e in ev:
    _lib.call('ofb_event_create', ctypes.byref(e))


def timed(fn, reps):
    fn()
    _lib.call('ofb_device_sync')
    _lib.call('ofb_event_record', ev[0], vp(None))
    for _ in range(reps):
        fn()
    _lib.call('ofb_event_record', ev[1], vp(None))
    _lib.call('ofb_event_sync', ev[1])
    ms = ctypes.c_float()
    _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(ms))
    return ms.value / reps


for name in sys.argv[1:] or ['mol26']:
    op = builders[name]()
    n = op.n_qubits
    dim = 1 << n
    plan = LinearQubitOperator(op, n).plan
    kmax = 4
    a = _lib.DeviceBuffer(dim * 16 * kmax)
    b = _lib.DeviceBuffer(dim * 16 * kmax)
    krylov.fill_random(dim * kmax, 1, False, a.ptr)
    reps = 2 if n >= 28 else 3
    single = timed(lambda: plan.apply_device(a.ptr, b.ptr), reps)
    rec = {'workload': name, 'n_qubits': n, 'terms': plan.n_terms, 'applied_terms': getattr(plan, 'n_reduced_terms', plan.n_terms),
           'single_column_ms': single}
    for k in (2, 4):
        ms = timed(lambda: plan.apply_device(a.ptr, b.ptr, k, dim), reps)
        rec['k%d_ms_per_column' % k] = ms / k
        rec['k%d_ratio_to_single' % k] = ms / k / single
        rec['k%d_term_amp_per_s' % k] = plan.n_terms * float(dim) * k / (ms / 1e3)
    print(json.dumps(rec), flush=True)
    a.free()
    b.free()
