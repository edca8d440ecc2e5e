#!/usr/bin/env python
"""Symmetry-sector H|psi> and ground-state timings.

Device-resident vectors, CUDA events on the launch stream; writes JSON lines.
    python scripts/sector_bench.py [hub44_half hub44_10 mol28 hub34_half] [--lanczos]
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import SectorLinearQubitOperator, get_ground_state, krylov  # noqa: E402

CASES = {
    'hub34_half': (lambda: hamiltonians.hubbard_jw(3, 4), 12, 0.0),
    'hub44_10': (lambda: hamiltonians.hubbard_jw(4, 4), 10, 0.0),
    'hub44_half': (lambda: hamiltonians.hubbard_jw(4, 4), 16, 0.0),
    'mol24': (lambda: hamiltonians.molecular_jw(12), 12, 0.0),
    'mol28': (lambda: hamiltonians.molecular_jw(14), 14, 0.0),
}


def event():
    e = ctypes.c_void_p()
    _lib.call('ofb_event_create', ctypes.byref(e))
    return e


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')] or ['hub34_half', 'hub44_10', 'hub44_half']
    lanczos = '--lanczos' in sys.argv
    tol = 1e-9
    for a in sys.argv[1:]:
        if a.startswith('--tol='):
            tol = float(a[6:])
    out_path = os.path.join('gpurun_out', 'sector_r01.jsonl')
    os.makedirs('gpurun_out', exist_ok=True)
    for name in args:
        build, ne, sz = CASES[name]
        t0 = time.perf_counter()
        op = build()
        n = op.n_qubits
        sec = SectorLinearQubitOperator(op, n, n_electrons=ne, sz_value=sz)
        plan, sector = sec.plan, sec.sector
        apply_plan = sec.apply_plan  # the orbit-reduced table under OFB_SUPPORT_HINTS=1, else `plan`
        _lib.call('ofb_device_sync')
        setup = time.perf_counter() - t0
        dim = sector.dim
        a = _lib.DeviceBuffer(dim * 16)
        b = _lib.DeviceBuffer(dim * 16)
        krylov.fill_random(dim, 1, False, a.ptr)
        for _ in range(3):
            sector.apply_device(apply_plan, a.ptr, b.ptr)
        e0, e1 = event(), event()
        reps = 5
        _lib.call('ofb_device_sync')
        _lib.call('ofb_event_record', e0, None)
        for _ in range(reps):
            sector.apply_device(apply_plan, a.ptr, b.ptr)
        _lib.call('ofb_event_record', e1, None)
        _lib.call('ofb_event_sync', e1)
        ms = ctypes.c_float()
        _lib.call('ofb_event_elapsed_ms', e0, e1, ctypes.byref(ms))
        ms = ms.value / reps
        herm = krylov.dotc(dim, a.ptr, b.ptr)
        line = {'case': name, 'n_qubits': n, 'n_electrons': ne, 'sz': sz, 'dim': dim,
                'full_dim': 1 << n, 'terms': plan.n_terms, 'applied_terms': apply_plan.n_terms, 'x_groups': plan.n_groups,
                'ms_per_apply': ms, 'setup_s': setup,
                'sector_term_amp_per_s': plan.n_terms * dim / (ms * 1e-3),
                'full_space_equivalent_term_amp_per_s': plan.n_terms * float(1 << n) / (ms * 1e-3),
                'algorithmic_GBps': 16.0 * dim * (plan.n_groups + 1) / (ms * 1e-3) / 1e9,
                'imag_of_vHv': herm.imag / max(abs(herm), 1e-300)}
        a.free()
        b.free()
        if lanczos:
            stats = {}
            t0 = time.perf_counter()
            dev_op = krylov.as_device_operator(sec)
            theta, vec, res = krylov.lanczos_lowest(dev_op, tol=tol, stats=stats)
            line.update({'lanczos_s': time.perf_counter() - t0, 'e0': theta, 'residual': res,
                         'matvecs': stats.get('matvecs')})
            vec.free()
        print(json.dumps(line), flush=True)
        with open(out_path, 'a') as f:
            f.write(json.dumps(line) + '\n')


if __name__ == '__main__':
    main()
