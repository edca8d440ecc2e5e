"""Drop-in subset of openfermion.linalg.sparse_tools backed by libofb200.

Mirrors the hot-path functions of /root/reference/src/openfermion/linalg/sparse_tools.py
(SURVEY.md section 8a rows a5-a12, a15): same names, arguments, return types and errors.
Materialisation runs the count/scan/fill CSC kernels (K4), the diagonal runs K3,
expectation runs the fused K2 kernel and get_ground_state / get_gap run the
device-resident Lanczos loop on top of K1.
"""
import itertools

import numpy
import scipy.sparse
import scipy.sparse.linalg

from openfermion_b200 import ops
from openfermion_b200.compiler import compile_fermion_terms, compile_qubit_operator
from openfermion_b200.linalg import krylov
from openfermion_b200.linalg.linear_qubit_operator import LinearQubitOperator
from openfermion_b200.ops import count_qubits
from openfermion_b200.plan import PauliPlan


def qubit_operator_sparse(qubit_operator, n_qubits=None):
    """scipy CSC of a QubitOperator (sparse_tools.py:136-194)."""
    if n_qubits is None:
        n_qubits = count_qubits(qubit_operator)
    if n_qubits < count_qubits(qubit_operator):
        raise ValueError('Invalid number of qubits specified.')
    x, z, c = compile_qubit_operator(qubit_operator, n_qubits)
    plan = PauliPlan(n_qubits, x, z, c)
    try:
        return plan.csc_host()
    finally:
        plan.close()


def jordan_wigner_sparse(fermion_operator, n_qubits=None):
    """scipy CSC of the Jordan-Wigner image of a FermionOperator (sparse_tools.py:76-133)."""
    if n_qubits is None:
        n_qubits = count_qubits(fermion_operator)
    rows = compile_fermion_terms(fermion_operator.terms, n_qubits)
    plan = PauliPlan.projected_rows(n_qubits, *rows)
    try:
        return plan.csc_host()
    finally:
        plan.close()


def _tensor_fermion_terms(operator):
    """Terms of a PolynomialTensor-like object in the reference's iteration order
    (ops/representations/polynomial_tensor.py __iter__, conversions.py:117-121): constant
    first, then each n-body tensor's non-zero entries in index-product order."""
    accumulated = ops.FermionOperator()
    tensors = operator.n_body_tensors

    def sort_key(key):  # () first, then by length, then the key digits read as an integer
        return (0, 0) if key == () else (len(key), int(''.join(map(str, key))))

    for key in sorted(tensors.keys(), key=sort_key):
        if key == ():
            accumulated += ops.FermionOperator((), tensors[key])
            continue
        tensor = numpy.asarray(tensors[key])
        n_modes = tensor.shape[0]
        for index in itertools.product(range(n_modes), repeat=len(key)):
            if tensor[index]:
                accumulated += ops.FermionOperator(tuple(zip(index, key)), tensor[index])
    return accumulated


def _diagonal_coulomb_fermion_terms(operator):
    """constant + sum_pq T_pq p^ q + sum_pq V_pq p^ p q^ q in the reference's order
    (transforms/opconversions/conversions.py:124-132)."""
    one_body = numpy.asarray(operator.one_body)
    two_body = numpy.asarray(operator.two_body)
    n = one_body.shape[0]
    out = ops.FermionOperator()
    out += ops.FermionOperator((), operator.constant)
    for p, q in itertools.product(range(n), repeat=2):
        out += ops.FermionOperator(((p, 1), (q, 0)), one_body[p, q])
        out += ops.FermionOperator(((p, 1), (p, 0), (q, 1), (q, 0)), two_body[p, q])
    return out


def _to_fermion_operator(operator):
    try:  # upstream converter when OpenFermion itself is installed
        from openfermion.transforms.opconversions import get_fermion_operator
        return get_fermion_operator(operator)
    except ImportError:
        pass
    if hasattr(operator, 'n_body_tensors'):
        return _tensor_fermion_terms(operator)
    if all(hasattr(operator, name) for name in ('one_body', 'two_body', 'constant')):
        return _diagonal_coulomb_fermion_terms(operator)
    raise TypeError('{} cannot be converted to FermionOperator'.format(type(operator)))


def get_sparse_operator(operator, n_qubits=None, trunc=None, hbar=1.0):
    """Type dispatch of sparse_tools.py:1249-1279.  Boson / quadrature operators are outside
    the Pauli-sum path and are passed to upstream if present."""
    names = {c.__name__ for c in type(operator).__mro__}
    if (names & {'DiagonalCoulombHamiltonian', 'PolynomialTensor'} or hasattr(operator, 'n_body_tensors')
            or all(hasattr(operator, name) for name in ('one_body', 'two_body', 'constant'))):
        return jordan_wigner_sparse(_to_fermion_operator(operator))
    if ops.is_fermion_operator(operator):
        return jordan_wigner_sparse(operator, n_qubits)
    if ops.is_qubit_operator(operator):
        return qubit_operator_sparse(operator, n_qubits)
    if names & {'BosonOperator', 'QuadOperator'}:
        from openfermion.linalg.sparse_tools import boson_operator_sparse  # upstream only
        return boson_operator_sparse(operator, trunc, hbar)
    raise TypeError('Failed to convert a {} to a sparse matrix.'.format(type(operator).__name__))


def get_linear_qubit_operator_diagonal(qubit_operator, n_qubits=None):
    """float64 diagonal of LinearQubitOperator(qubit_operator) (sparse_tools.py:197-247)."""
    if n_qubits is None:
        n_qubits = count_qubits(qubit_operator)
    if n_qubits < count_qubits(qubit_operator):
        raise ValueError('Invalid number of qubits specified.')
    # The reference accumulates into a float64 array and numpy refuses to cast a complex
    # coefficient into it; only X/Y-free strings reach
    # that accumulation.
    for term, coefficient in qubit_operator.terms.items():
        if isinstance(coefficient, (complex, numpy.complexfloating)):
            if all(action == 'Z' for _, action in term):
                raise TypeError(
                    "Cannot cast a complex coefficient into the float64 diagonal "
                    "(call .compress() on the operator first)."
                )
    x, z, c = compile_qubit_operator(qubit_operator, n_qubits)
    plan = PauliPlan(n_qubits, x, z, c)
    try:
        return plan.diagonal_host()
    finally:
        plan.close()


def _is_sparse(obj):
    return scipy.sparse.issparse(obj)


def expectation(operator, state):
    """<state|operator|state> (sparse_tools.py:629-671)."""
    if _is_sparse(state):
        # density matrix branch: not a Pauli-sum-on-state-vector path; evaluated with scipy
        # exactly as the reference does (DESIGN.md "out of scope")
        if isinstance(operator, scipy.sparse.linalg.LinearOperator):
            raise ValueError(
                'Taking the expectation of a LinearOperator with '
                'a density matrix is not supported.'
            )
        product = state * operator
        return numpy.sum(product.diagonal())
    if isinstance(state, numpy.ndarray):
        if isinstance(operator, LinearQubitOperator):
            if state.size != operator.shape[1]:
                raise ValueError('Dimension mismatch')
            return operator.plan.expectation_host(state)
        applied = krylov.apply_any(operator, state)
        if len(state.shape) == 1:
            return numpy.dot(numpy.conjugate(state), applied)
        return numpy.dot(numpy.conjugate(state.T), applied)[0, 0]
    raise ValueError('Input state must be a numpy array or a sparse matrix.')


def variance(operator, state):
    """<H^2> - <H>^2 (sparse_tools.py:674-690)."""
    return expectation(operator**2, state) - expectation(operator, state) ** 2


def get_ground_state(sparse_operator, initial_guess=None):
    """Lowest eigenpair (sparse_tools.py:566-587: eigsh(k=1, which='SA')), computed by the
    device-resident Lanczos loop.  Returns (float eigenvalue, complex128[dim] eigenstate)."""
    values, vectors = krylov.lowest_eigenpairs(sparse_operator, 1, initial_guess)
    return values[0], vectors[:, 0]


def get_gap(sparse_operator, initial_guess=None):
    """|lambda_1 - lambda_0| (sparse_tools.py:1078-1095)."""
    if _is_sparse(sparse_operator):
        diff = sparse_operator - sparse_operator.getH()
        if diff.nnz and numpy.max(numpy.abs(diff.data)) > 1e-8:
            raise ValueError('sparse_operator must be Hermitian.')
    values, _ = krylov.lowest_eigenpairs(sparse_operator, 2, initial_guess)
    return abs(values[1] - values[0])


def inner_product(state_1, state_2):
    """conj(state_1) . state_2 (sparse_tools.py:1098-1100)."""
    return numpy.dot(state_1.conjugate(), state_2)


# ----------------------------------------------------------------------------
# Jordan-Wigner symmetry-sector helpers
# (sparse_tools.py:273-345, 375-423, 477-513).  Index selection is host arithmetic as in
# the reference; the eigen-solve inside the sector runs on the device (CSR SpMV + Lanczos).
# ----------------------------------------------------------------------------
def jw_number_indices(n_electrons, n_qubits):
    """Indices of all basis states with n_electrons occupied modes, in the reference's order
    (itertools.combinations over bit positions, sparse_tools.py:273-292)."""
    occupations = itertools.combinations(range(n_qubits), n_electrons)
    return [sum(1 << n for n in occupation) for occupation in occupations]


def jw_sz_indices(sz_value, n_qubits, n_electrons=None, up_index=None, down_index=None):
    """Indices of basis states with fixed S_z (sparse_tools.py:295-372); up = 2i, down = 2i+1
    by default (utils/indexing.py:17,29)."""
    up_index = up_index or (lambda i: 2 * i)
    down_index = down_index or (lambda i: 2 * i + 1)
    if n_qubits % 2 != 0:
        raise ValueError('Number of qubits must be even')
    if not (2.0 * sz_value).is_integer():
        raise ValueError('Sz value must be an integer or half-integer')
    n_sites = n_qubits // 2
    sz_integer = int(2.0 * sz_value)
    indices = []

    def pair(outer_map, n_outer, inner_map, n_inner):
        # every arrangement of the outer spin species with every arrangement of the inner one
        inner = list(itertools.combinations(range(n_sites), n_inner))
        for outer_occ in itertools.combinations(range(n_sites), n_outer):
            outer_modes = [outer_map(i) for i in outer_occ]
            for inner_occ in inner:
                occupation = outer_modes + [inner_map(i) for i in inner_occ]
                indices.append(sum(1 << (n_qubits - 1 - m) for m in occupation))

    if n_electrons is not None:
        if (n_electrons + sz_integer) % 2 != 0 or n_electrons < abs(sz_integer):
            raise ValueError('The specified particle number and sz value are incompatible.')
        num_up = (n_electrons + sz_integer) // 2
        pair(up_index, num_up, down_index, n_electrons - num_up)
    else:
        more_map, less_map = (down_index, up_index) if sz_integer < 0 else (up_index, down_index)
        for n in range(abs(sz_integer), n_sites + 1):
            pair(more_map, n, less_map, n - abs(sz_integer))
    return indices


def jw_number_restrict_operator(operator, n_electrons, n_qubits=None):
    if n_qubits is None:
        n_qubits = int(numpy.log2(operator.shape[0]))
    select = jw_number_indices(n_electrons, n_qubits)
    return operator[numpy.ix_(select, select)]


def jw_sz_restrict_operator(operator, sz_value, n_electrons=None, n_qubits=None, up_index=None,
                            down_index=None):
    if n_qubits is None:
        n_qubits = int(numpy.log2(operator.shape[0]))
    select = jw_sz_indices(sz_value, n_qubits, n_electrons=n_electrons, up_index=up_index,
                           down_index=down_index)
    return operator[numpy.ix_(select, select)]


def jw_get_ground_state_at_particle_number(sparse_operator, particle_number):
    """Lowest eigenpair inside a particle-number sector, expanded back to 2^n amplitudes
    (sparse_tools.py:477-513)."""
    n_qubits = int(numpy.log2(sparse_operator.shape[0]))
    restricted = jw_number_restrict_operator(sparse_operator, particle_number, n_qubits)
    if restricted.shape[0] - 1 <= 1:
        eigvals, eigvecs = numpy.linalg.eigh(restricted.toarray())
        energy, state = eigvals[0], eigvecs[:, 0]
    else:
        energy, state = get_ground_state(scipy.sparse.csc_matrix(restricted))
    expanded = numpy.zeros(2**n_qubits, dtype=complex)
    expanded[jw_number_indices(particle_number, n_qubits)] = state
    return energy, expanded
