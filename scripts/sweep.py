#!/usr/bin/env python
"""H|psi> throughput sweep of BASELINE.json config 4 (synthetic molecular Hamiltonians,
n = 20..28) plus the Hubbard lattices, device-resident, CUDA-event timed.  One JSON line each."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import LinearQubitOperator, krylov  # noqa: E402

PEAK = 6536.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
cases = [('molecular', n, (lambda n=n: hamiltonians.molecular_jw(n // 2, seed=1234))) for n in (20, 22, 24, 26, 28)]
cases += [('hubbard 3x4', 24, lambda: hamiltonians.hubbard_jw(3, 4)), ('hubbard 2x7', 28, lambda: hamiltonians.hubbard_jw(2, 7)),
          ('hubbard 3x5', 30, lambda: hamiltonians.hubbard_jw(3, 5)), ('hubbard 4x4', 32, lambda: hamiltonians.hubbard_jw(4, 4))]
vp = ctypes.c_void_p
for name, n, build in cases:
    op = build()
    dim = 1 << n
    plan = LinearQubitOperator(op, n).plan
    a, b = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, 1, False, a.ptr)
    ev = [vp(), vp()]
    for e in ev:
        _lib.call('ofb_event_create', ctypes.byref(e))
    for _ in range(3):
        plan.apply_device(a.ptr, b.ptr)
    reps = 5 if n <= 26 else 3
    _lib.call('ofb_event_record', ev[0], vp(None))
    for _ in range(reps):
        plan.apply_device(a.ptr, b.ptr)
    _lib.call('ofb_event_record', ev[1], vp(None))
    _lib.call('ofb_event_sync', ev[1])
    ms = ctypes.c_float()
    _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(ms))
    sec = ms.value / 1e3 / reps
    alg = 16.0 * dim * (plan.n_groups + 1)
    print(json.dumps({'workload': name, 'n_qubits': n, 'terms': plan.n_terms, 'x_groups': plan.n_groups,
                      'ms_per_matvec': sec * 1e3, 'term_amp_per_s': plan.n_terms * dim / sec,
                      'algorithmic_GBs': alg / sec / 1e9, 'frac_of_measured_peak': alg / sec / 1e9 / PEAK}), flush=True)
    a.free()
    b.free()
    plan.close()
