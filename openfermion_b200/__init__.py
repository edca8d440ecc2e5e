"""B200-native (sm_100a) engine for OpenFermion's Pauli-sum numeric hot path.

`openfermion_b200.linalg` mirrors the hot-path subset of `openfermion.linalg`; the
kernels live in libofb200.so (C ABI: include/ofb200.h).  There is no CPU fallback.
"""
from openfermion_b200.ops import FermionOperator, QubitOperator, count_qubits  # noqa: F401

__version__ = '0.1.0'
