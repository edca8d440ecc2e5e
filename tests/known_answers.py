"""Known-answer vectors restated from the reference's own tests for the hot path.

Sources (under /root/reference/src/openfermion/linalg):
  MATVEC   linear_qubit_operator_test.py:99-176
  CSC      sparse_tools_test.py:1774-1842 (exact .data and .indices lists)
  DIAGONAL sparse_tools_test.py:194-244
Each entry builds its operator with openfermion_b200.ops.QubitOperator so that the same
table pins both the CPU oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).
"""
import numpy

from openfermion_b200.ops import QubitOperator as Q


def _multi():
    return Q(()) + 2 * Q('Y2') + Q(((0, 'Z'), (1, 'X')), 10.0)


_v8 = numpy.array([1, 2, 3, 4, 5, 6, 7, 8])

# (name, operator, n_qubits or None, vec, expected)
MATVEC = [
    ('zero', Q(), 3, _v8, numpy.zeros(8)),
    ('x1', Q('X1'), None, numpy.array([1, 2, 3, 4]), numpy.array([2, 1, 4, 3])),
    ('y1', Q('Y1'), None, numpy.array([1, 2, 3, 4], dtype=complex),
     1.0j * numpy.array([-2, 1, -4, 3], dtype=complex)),
    ('z0', Q('Z0'), 2, numpy.array([1, 2, 3, 4]), numpy.array([1, 2, -3, -4])),
    ('z3', Q('Z3'), None, numpy.arange(1, 17),
     numpy.array([1, -2, 3, -4, 5, -6, 7, -8, 9, -10, 11, -12, 13, -14, 15, -16])),
    ('z0x1', Q('Z0 X1'), None, numpy.array([1, 2, 3, 4]), numpy.array([2, 1, -4, -3])),
    ('multiple_terms', _multi(), None, _v8,
     10 * numpy.array([3, 4, 1, 2, -7, -8, -5, -6])
     + 2j * numpy.array([-2, 1, -4, 3, -6, 5, -8, 7]) + _v8),
]

# (name, operator, n_qubits or None, data list, indices list or None, shape)
CSC = [
    ('Y', Q(((0, 'Y'),)), None, [1j, -1j], [1, 0], (2, 2)),
    ('ZX', Q(((0, 'Z'), (1, 'X')), 2.0), None, [2.0, 2.0, -2.0, -2.0], [1, 0, 3, 2], (4, 4)),
    ('ZIZ', Q(((0, 'Z'), (2, 'Z'))), None, [1, -1, 1, -1, -1, 1, -1, 1], list(range(8)), (8, 8)),
    ('combo', Q(((0, 'Y'), (1, 'X')), -0.1j) + Q(((0, 'X'), (1, 'Z')), 3.0 + 2.0j), None,
     [3 + 2j, 0.1, 0.1, -3 - 2j, 3 + 2j, -0.1, -0.1, -3 - 2j], [2, 3, 2, 3, 0, 1, 0, 1], (4, 4)),
    ('zero_1q', Q((), 0.0), 1, [], [], (2, 2)),
    ('zero_5q', Q((), 0.0), 5, [], [], (32, 32)),
    ('identity_1q', Q(()), 1, [1] * 2, [0, 1], (2, 2)),
    ('identity_5q', Q(()), 5, [1] * 32, list(range(32)), (32, 32)),
    ('linearity', Q(()) + Q(tuple((i, 'Z') for i in range(4)), 1.0), None, [2] * 8,
     [0, 3, 5, 6, 9, 10, 12, 15], (16, 16)),
]

# (name, operator, n_qubits or None, expected diagonal)
DIAGONAL = [
    ('zero', Q(), 3, numpy.zeros(8)),
    ('x0y1', Q('X0 Y1'), None, numpy.zeros(4)),
    ('z0z2', Q('Z0 Z2'), None, numpy.array([1, -1, 1, -1, -1, 1, -1, 1])),
]
