#!/usr/bin/env python
"""Where the milliseconds of config 1 (LiH, 12 qubits) go: get_sparse_operator and get_ground_state
split into host stages; repeated calls in one process (development aid)."""
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import get_ground_state, get_sparse_operator, krylov  # noqa: E402


def main():
    _lib.require_device()
    data = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'lih_sto3g.npz'))
    ham = hamiltonians.molecular_hamiltonian(float(data['nuclear_repulsion']), data['one_body_integrals'],
                                             data['two_body_integrals'])
    out = {'sparse_s': [], 'ground_state_s': [], 'csr_upload_s': [], 'lanczos_s': [], 'download_s': [],
           'launches': [], 'matvecs': []}
    for _ in range(5):
        _lib.call('ofb_device_sync')
        t = time.perf_counter()
        csc = get_sparse_operator(ham)
        out['sparse_s'].append(time.perf_counter() - t)
        t = time.perf_counter()
        e0, psi = get_ground_state(csc)
        out['ground_state_s'].append(time.perf_counter() - t)
        # the same call in stages
        t = time.perf_counter()
        dev = krylov.as_device_operator(csc)
        _lib.call('ofb_device_sync')
        out['csr_upload_s'].append(time.perf_counter() - t)
        launches = _lib.launch_count()
        stats = {}
        t = time.perf_counter()
        energy, vec, res = krylov.lanczos_lowest(dev, stats=stats)
        out['lanczos_s'].append(time.perf_counter() - t)
        out['launches'].append(_lib.launch_count() - launches)
        out['matvecs'].append(stats.get('matvecs'))
        t = time.perf_counter()
        krylov.download(dev.dim, vec.ptr)
        out['download_s'].append(time.perf_counter() - t)
        vec.free()
    out['E0'] = e0
    print(json.dumps(out))


if __name__ == '__main__':
    main()
