#!/usr/bin/env python
"""QubitDavidson behind the C ABI (ofb_davidson_lowest_n) against the host-driven loop
(OFB_DAVIDSON_HOST=1) on the synthetic molecular Hamiltonian: wall time, iterations, eigenvalue.

    python scripts/davidson_timing.py [--qubits 12,16,20,24]"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import DavidsonOptions, QubitDavidson  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qubits', default='12,16,20,24')
    ap.add_argument('--max-subspace', type=int, default=16)
    ap.add_argument('--eps', type=float, default=1e-6)
    args = ap.parse_args()
    _lib.require_device()
    for n in [int(t) for t in args.qubits.split(',')]:
        op = hamiltonians.molecular_jw(n // 2, seed=1234)
        options = DavidsonOptions(max_subspace=args.max_subspace, eps=args.eps, max_iterations=600)
        dav = QubitDavidson(op, n, options)
        diag = numpy.real(dav.linear_operator_diagonal)
        # lowest-diagonal determinant with n/2 electrons
        idx = numpy.arange(1 << n, dtype=numpy.int64)
        pop = numpy.zeros(1 << n, dtype=numpy.int8)
        for b in range(n):
            pop += ((idx >> b) & 1).astype(numpy.int8)
        members = idx[pop == n // 2]
        start = int(members[int(numpy.argmin(diag[members]))])
        guess = numpy.zeros((1 << n, 1), dtype=complex)
        guess[start, 0] = 1.0
        rec = {'n_qubits': n, 'terms': len(op), 'max_subspace': args.max_subspace, 'eps': args.eps}
        for label, env in (('native', None), ('host_loop', '1'), ('native_again', None)):
            if env:
                os.environ['OFB_DAVIDSON_HOST'] = env
            else:
                os.environ.pop('OFB_DAVIDSON_HOST', None)
            launches = _lib.launch_count()
            t0 = time.perf_counter()
            ok, values, _ = dav.get_lowest_n(1, guess)
            rec[label] = {'s': time.perf_counter() - t0, 'success': bool(ok), 'eigenvalue': float(values[0]),
                          'launches': _lib.launch_count() - launches,
                          'iterations': getattr(dav, 'last_iterations', None) if not env else None}
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
