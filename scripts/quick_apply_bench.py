#!/usr/bin/env python
"""Quick device-resident H|psi> timing of a few workloads (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, hamiltonians
from openfermion_b200.linalg import LinearQubitOperator, krylov

which = sys.argv[1:] or ['mol26', 'hub28']
builders = {'mol24': lambda: hamiltonians.molecular_jw(12), 'mol26': lambda: hamiltonians.molecular_jw(13),
            'mol28': lambda: hamiltonians.molecular_jw(14), 'hub24': lambda: hamiltonians.hubbard_jw(3, 4),
            'hub28': lambda: hamiltonians.hubbard_jw(2, 7), 'hub30': lambda: hamiltonians.hubbard_jw(3, 5),
            'rand24c': lambda: hamiltonians.random_pauli_sum(24, 400, 1, True),
            'bk24': lambda: hamiltonians.molecular_jw(12, encoding='bk'), 'bk28': lambda: hamiltonians.molecular_jw(14, encoding='bk'),
            'hubbk28': lambda: hamiltonians.hubbard_jw(2, 7, encoding='bk')}
for name in which:
    op = builders[name]()
    n = op.n_qubits
    dim = 1 << n
    plan = LinearQubitOperator(op, n).plan
    a = _lib.DeviceBuffer(dim * 16)
    b = _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, 1, False, a.ptr)
    for _ in range(2):
        plan.apply_device(a.ptr, b.ptr)
    _lib.call('ofb_device_sync')
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        plan.apply_device(a.ptr, b.ptr)
    _lib.call('ofb_device_sync')
    dt = (time.perf_counter() - t0) / reps
    print(name, 'T', plan.n_terms, 'G', plan.n_groups, 'ms %.2f' % (dt * 1e3),
          'term*amp/s %.3e' % (plan.n_terms * dim / dt),
          'alg GB/s %.0f' % (16.0 * dim * (plan.n_groups + 1) / dt / 1e9),
          'check', krylov.dotc(dim, a.ptr, b.ptr))
