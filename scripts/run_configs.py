#!/usr/bin/env python
"""Runs the BASELINE.json configurations 1-4 end to end on one B200 and prints one JSON
object with timings and self-checks (used for DESIGN.md section 6; not the bench contract).

    python scripts/run_configs.py [--skip 2,4]
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import (DavidsonOptions, LinearQubitOperator, QubitDavidson,  # noqa: E402
                                     expectation, get_ground_state, get_sparse_operator,
                                     qubit_operator_sparse)
from openfermion_b200.linalg import krylov  # noqa: E402


def timed(fn):
    _lib.call('ofb_device_sync')
    t0 = time.perf_counter()
    out = fn()
    _lib.call('ofb_device_sync')
    return out, time.perf_counter() - t0


LIH_HDF5 = os.environ.get(
    'OFB_LIH_HDF5', '/root/reference/src/openfermion/testing/data/H1-Li1_sto-3g_singlet_1.45.hdf5')


def config1():
    data = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'lih_sto3g.npz'))
    if os.path.exists(LIH_HDF5):
        # the bundled molecule file itself, through the pure-Python HDF5 reader
        from openfermion_b200 import hdf5
        ham = hdf5.load_molecular_hamiltonian(LIH_HDF5)
        source = 'hdf5:' + os.path.basename(LIH_HDF5)
    else:  # the GPU box has no reference tree: same integrals from the committed fixture
        ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                                 data['two_body_integrals'])
        source = 'tests/golden/lih_sto3g.npz'
    csc, t_csc = timed(lambda: get_sparse_operator(ham))
    (e0, psi), t_gs = timed(lambda: get_ground_state(csc))
    return {'config': 'LiH sto-3g 12q: get_sparse_operator + get_ground_state', 'integrals': source,
            'nnz': int(csc.nnz), 'sparse_s': t_csc, 'ground_state_s': t_gs, 'E0': e0,
            'abs_err_vs_reference_eigsh': abs(e0 - float(data['reference_e0'])),
            'abs_err_vs_file_fci': abs(e0 - float(data['fci_energy']))}


def config2():
    op = hamiltonians.molecular_jw(10, seed=1234, dyadic_bits=16)
    n = 20
    csc, t_csc_first = timed(lambda: qubit_operator_sparse(op, n))  # incl. one-time module load / staging buffers
    del csc
    csc, t_csc = timed(lambda: qubit_operator_sparse(op, n))
    rng = numpy.random.default_rng(0)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= numpy.linalg.norm(psi)
    lin = LinearQubitOperator(op, n)
    e_free, t_exp = timed(lambda: expectation(lin, psi))
    # sampled columns of the CSC against the matrix-free operator applied to basis vectors
    cols = [0, 1, 12345, 2**19 + 7, 2**20 - 1]
    worst = 0.0
    for c in cols:
        basis = numpy.zeros(2**n, dtype=complex)
        basis[c] = 1.0
        want = lin * basis
        got = numpy.zeros(2**n, dtype=complex)
        lo, hi = csc.indptr[c], csc.indptr[c + 1]
        got[csc.indices[lo:hi]] = csc.data[lo:hi]
        worst = max(worst, float(numpy.max(numpy.abs(got - want))))
        assert numpy.all(numpy.diff(csc.indices[lo:hi]) > 0)
    e_csc = numpy.vdot(psi, csc @ psi)
    return {'config': 'synthetic molecular 20q (dyadic integrals): CSC build + expectation',
            'terms': len(op), 'nnz': int(csc.nnz), 'index_dtype': str(csc.indices.dtype),
            'csc_build_s': t_csc, 'csc_build_first_call_s': t_csc_first, 'expectation_s': t_exp, 'expectation': [e_free.real, e_free.imag],
            'csc_vs_matrix_free_expectation_abs': abs(e_csc - e_free),
            'sampled_columns_max_abs_diff': worst}


def config3():
    op = hamiltonians.hubbard_jw(3, 4)
    lin = LinearQubitOperator(op, 24)
    stats = {}
    dev = krylov.as_device_operator(lin)
    (theta, vec, resid), t = timed(lambda: krylov.lanczos_lowest(dev, None, (), 1e-10, stats=stats))
    d_in = _lib.DeviceBuffer(2**24 * 16)
    d_out = _lib.DeviceBuffer(2**24 * 16)
    krylov.fill_random(2**24, 1, False, d_in.ptr)
    lin.plan.apply_device(d_in.ptr, d_out.ptr)
    _, t_mv = timed(lambda: [lin.plan.apply_device(d_in.ptr, d_out.ptr) for _ in range(10)])
    return {'config': 'Fermi-Hubbard 3x4 (24q) matrix-free Lanczos ground state',
            'terms': lin.plan.n_terms, 'x_groups': lin.plan.n_groups, 'E0': theta, 'residual': resid,
            'lanczos_s': t, 'matvecs': stats.get('matvecs'), 'restarts': stats.get('restarts'),
            'matvec_ms': 1e3 * t_mv / 10,
            'matvec_algorithmic_GBs': 16.0 * 2**24 * (lin.plan.n_groups + 1) / (t_mv / 10) / 1e9}


def config4(n=28, iterations=3):
    op = hamiltonians.molecular_jw(n // 2, seed=1234)
    numpy.random.seed(7)
    dav, t_setup = timed(lambda: QubitDavidson(op, n, DavidsonOptions(max_subspace=8, eps=1e-6)))
    guess = numpy.zeros((2**n, 1), dtype=complex)
    diag = dav.linear_operator_diagonal
    guess[int(numpy.argmin(diag)), 0] = 1.0  # lowest diagonal entry: the usual Davidson start
    (ok, vals, vecs), t = timed(lambda: dav.get_lowest_n(1, guess, max_iterations=iterations))
    return {'config': 'synthetic molecular %dq: QubitDavidson, max_subspace=8, %d iterations' % (n, iterations),
            'setup_s_(plan+diagonal)': t_setup, 'davidson_s': t, 'converged': bool(ok),
            'ritz_value': float(vals[0]), 'lowest_diagonal': float(diag.min())}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip', default='')
    ap.add_argument('--davidson-qubits', type=int, default=28)
    args = ap.parse_args()
    skip = {s for s in args.skip.split(',') if s}
    _lib.require_device()
    out = {}
    for name, fn in (('1', config1), ('2', config2), ('3', config3),
                     ('4', lambda: config4(args.davidson_qubits))):
        if name in skip:
            continue
        out['config%s' % name] = fn()
        print(json.dumps({name: out['config%s' % name]}), flush=True)
    print(json.dumps(out))
