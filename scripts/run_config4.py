#!/usr/bin/env python
"""BASELINE.json config 4, the eigensolve half: QubitDavidson (linalg/davidson.py:99-221 mirrored
in openfermion_b200/linalg/davidson.py) on the synthetic molecular-type Hamiltonian in the FULL
2^n space, converged to the reference's criterion (max residual < eps, success=True), and checked
against an independent solve: Lanczos inside the symmetry sector of the start determinant (K7).

    python scripts/run_config4.py --qubits 28 [--max-subspace 16] [--eps 1e-6] [--max-iterations 400]

The start vector is the lowest-diagonal determinant of the (n/2 electrons, S_z = 0) sector; H and
the diagonal preconditioner conserve particle number and S_z, so the Davidson iterates stay in
that sector although they are stored and multiplied as full 2^n vectors."""
import argparse
import json
import logging
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openfermion_b200 import _lib, hamiltonians, sectors  # noqa: E402
from openfermion_b200.linalg import (DavidsonOptions, QubitDavidson, SectorLinearQubitOperator,  # noqa: E402
                                     krylov)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qubits', type=int, default=28)
    ap.add_argument('--max-subspace', type=int, default=16)
    ap.add_argument('--eps', type=float, default=1e-6)
    ap.add_argument('--max-iterations', type=int, default=400)
    ap.add_argument('--skip-sector', action='store_true')
    ap.add_argument('--native', action='store_true',
                    help='no INFO logging: the solver runs behind the C ABI (ofb_davidson_lowest_n)')
    args = ap.parse_args()
    if not args.native:  # the reference's per-iteration log line needs the host-driven loop
        logging.basicConfig(level=logging.INFO, format='%(message)s', stream=sys.stderr)
    _lib.require_device()
    n = args.qubits
    op = hamiltonians.molecular_jw(n // 2, seed=1234)
    out = {'config': 'synthetic molecular %dq: QubitDavidson in the full space, max_subspace=%d, eps=%g'
                     % (n, args.max_subspace, args.eps), 'terms': len(op)}
    t0 = time.perf_counter()
    dav = QubitDavidson(op, n, DavidsonOptions(max_subspace=args.max_subspace, eps=args.eps,
                                               max_iterations=args.max_iterations))
    diag = dav.linear_operator_diagonal
    sector_op = SectorLinearQubitOperator(dav.linear_operator, n_electrons=n // 2, sz_value=0.0)
    members = numpy.asarray(sector_op.indices(), dtype=numpy.int64)  # jw_sz_indices, built on the device
    start = int(members[int(numpy.argmin(diag[members]))])
    guess = numpy.zeros((1 << n, 1), dtype=complex)
    guess[start, 0] = 1.0
    out['setup_s'] = time.perf_counter() - t0
    out['start_determinant'] = start
    out['start_diagonal'] = float(diag[start])
    launches = _lib.launch_count()
    t0 = time.perf_counter()
    success, values, vectors = dav.get_lowest_n(1, guess)
    out['davidson_s'] = time.perf_counter() - t0
    out['success'] = bool(success)
    out['eigenvalue'] = float(values[0])
    out['solver'] = 'ofb_davidson_lowest_n' if args.native else 'host-driven loop'
    out['iterations'] = getattr(dav, 'last_iterations', None)
    out['kernel_launches'] = _lib.launch_count() - launches
    vec = vectors[:, 0]
    out['norm'] = float(numpy.linalg.norm(vec))
    out['weight_outside_sector'] = float(1.0 - numpy.sum(numpy.abs(vec[members]) ** 2) / numpy.vdot(vec, vec).real)
    del vectors, guess
    print(json.dumps(out), flush=True)
    if not args.skip_sector:
        # independent check: Lanczos on the sector kernel (different kernel, different solver)
        stats = {}
        t0 = time.perf_counter()
        energy, _, residual = krylov.lanczos_lowest(krylov.as_device_operator(sector_op), stats=stats)
        out['sector_lanczos'] = {'E0': energy, 'residual': residual, 'matvecs': stats.get('matvecs'),
                                 'seconds': time.perf_counter() - t0, 'dim': int(sector_op.shape[0])}
        out['relative_difference'] = abs(out['eigenvalue'] - energy) / abs(energy)
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
