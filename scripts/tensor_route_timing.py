#!/usr/bin/env python
"""get_sparse_operator on InteractionOperator-like tensors (the reference's molecular route:
get_fermion_operator + jordan_wigner_sparse): rows built array-wise from the tensors against the
term-by-term route (OFB_TENSOR_ROWS=0), wall time per call and byte equality."""
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib, hamiltonians  # noqa: E402
from openfermion_b200.linalg import get_sparse_operator  # noqa: E402

_lib.require_device()
for n_orbitals in [int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else '6,7,8,9').split(',')]:
    constant, one, two = hamiltonians.random_interaction_tensors(n_orbitals, seed=1234)
    ham = hamiltonians.InteractionTensor(constant, one, two)
    rec = {'n_qubits': 2 * n_orbitals}
    mats = {}
    for label, env in (('tensor_rows', '1'), ('term_by_term', '0'), ('tensor_rows_again', '1')):
        os.environ['OFB_TENSOR_ROWS'] = env
        _lib.call('ofb_device_sync')
        t0 = time.perf_counter()
        mats[label] = get_sparse_operator(ham)
        rec[label + '_s'] = time.perf_counter() - t0
    a, b = mats['tensor_rows'], mats['term_by_term']
    rec['nnz'] = int(a.nnz)
    rec['identical'] = bool(numpy.array_equal(a.indptr, b.indptr) and numpy.array_equal(a.indices, b.indices)
                            and a.data.tobytes() == b.data.tobytes())
    print(json.dumps(rec), flush=True)
