#!/usr/bin/env python
"""One-GPU lattice path: H|psi> with the index bits in their natural order against the
locality-ordered relabelling (shard_basis.locality_order with zero shard bits), including the
two relabelling passes (ofb_vec_relabel) that keep input and output in the reference's order.
   python scripts/local_order_bench.py hub24 hub30 hub32"""
import ctypes
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openfermion_b200 import _lib, compiler, hamiltonians  # noqa: E402
from openfermion_b200.linalg import krylov  # noqa: E402
from openfermion_b200.plan import PauliPlan  # noqa: E402
from openfermion_b200.shard_basis import ShardBasis  # noqa: E402

builders = {'hub24': lambda: hamiltonians.hubbard_jw(3, 4), 'hub28': lambda: hamiltonians.hubbard_jw(2, 7),
            'hub30': lambda: hamiltonians.hubbard_jw(3, 5), 'hub32': lambda: hamiltonians.hubbard_jw(4, 4),
            'hubbk28': lambda: hamiltonians.hubbard_jw(2, 7, encoding='bk'),
            'mol24': lambda: hamiltonians.molecular_jw(12)}
vp = ctypes.c_void_p
ev = [vp(), vp()]
for e in ev:
    _lib.call('ofb_event_create', ctypes.byref(e))


def timed(fn, reps=3):
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


def cols_table(values):
    t = numpy.zeros(40, dtype=numpy.uint64)
    t[:len(values)] = numpy.array(values, dtype=numpy.uint64)
    return t


for name in sys.argv[1:] or ['hub24']:
    op = builders[name]()
    n = op.n_qubits
    dim = 1 << n
    x, z, c = compiler.compile_qubit_operator(op, n)
    basis = ShardBasis.for_sharding(n, 0, x, local_order='locality')
    rec = {'workload': name, 'n_qubits': n, 'terms': len(c), 'relabelled': not basis.identity}
    natural = PauliPlan(n, x, z, c)
    a, b = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, 1, False, a.ptr)
    rec['natural_ms'] = timed(lambda: natural.apply_device(a.ptr, b.ptr))
    if not basis.identity:
        xb, zb, cb = basis.transform_terms(x, z, c)
        plan = PauliPlan(n, xb, zb, cb)
        to_basis = cols_table(basis.cols)  # psi'[c'] = psi[B^-1 c']
        to_natural = cols_table([int(v) for v in basis.to_basis_index(numpy.array([1 << k for k in range(n)], dtype=numpy.uint64))])
        want_head = krylov.download(4096, b.ptr)
        e_nat = krylov.dotc(dim, a.ptr, b.ptr)
        # b <- relabel(a); a' := b; apply; relabel back
        t_in = timed(lambda: _lib.call('ofb_vec_relabel', ctypes.c_int(n), _lib.ptr(to_basis), vp(a.ptr), vp(b.ptr), None))
        natural.close()
        rec['relabel_pass_ms'] = t_in
        if n >= 32:  # three 64 GiB vectors do not fit: time only (the input is overwritten)
            rec['relabelled_ms'] = timed(lambda: plan.apply_device(b.ptr, a.ptr))
        else:
            w = _lib.DeviceBuffer(dim * 16)
            rec['relabelled_ms'] = timed(lambda: plan.apply_device(b.ptr, w.ptr))
            _lib.call('ofb_vec_relabel', ctypes.c_int(n), _lib.ptr(to_natural), vp(w.ptr), vp(b.ptr), None)
            got_head = krylov.download(4096, b.ptr)
            e_rel = krylov.dotc(dim, a.ptr, b.ptr)
            rec['max_abs_diff_head'] = float(numpy.max(numpy.abs(got_head - want_head)))
            rec['vHv_natural'] = [e_nat.real, e_nat.imag]
            rec['vHv_relabelled'] = [e_rel.real, e_rel.imag]
            w.free()
        rec['speedup_kernel_only'] = rec['natural_ms'] / rec['relabelled_ms']
        rec['speedup_with_both_relabel_passes'] = rec['natural_ms'] / (rec['relabelled_ms'] + 2 * t_in)
        plan.close()
    print(json.dumps(rec), flush=True)
    a.free()
    b.free()
