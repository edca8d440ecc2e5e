"""Drop-in for the hot-path subset of `openfermion.linalg` (linalg/__init__.py:12-85)."""
from openfermion_b200.linalg.linear_qubit_operator import (
    LinearQubitOperator,
    LinearQubitOperatorOptions,
    ParallelLinearQubitOperator,
    SectorLinearQubitOperator,
    apply_operator,
    generate_linear_qubit_operator,
)
from openfermion_b200.linalg.sparse_tools import (
    expectation,
    get_gap,
    get_ground_state,
    get_linear_qubit_operator_diagonal,
    get_number_preserving_sparse_operator,
    get_sparse_operator,
    inner_product,
    jordan_wigner_sparse,
    jw_get_ground_state_at_particle_number,
    jw_number_indices,
    jw_number_restrict_operator,
    jw_number_restrict_state,
    jw_sz_indices,
    jw_sz_restrict_operator,
    jw_sz_restrict_state,
    qubit_operator_sparse,
    variance,
)
from openfermion_b200.linalg.davidson import (
    Davidson,
    DavidsonError,
    DavidsonOptions,
    QubitDavidson,
    SparseDavidson,
    append_random_vectors,
    generate_random_vectors,
    orthonormalize,
)
