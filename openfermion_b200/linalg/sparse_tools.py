"""Drop-in subset of openfermion.linalg.sparse_tools backed by libofb200.

Mirrors the hot-path functions of /root/reference/src/openfermion/linalg/sparse_tools.py
(SURVEY.md section 8a rows a5-a12, a15): same names, arguments, return types and errors.
Materialisation runs the count/scan/fill CSC kernels (K4), the diagonal runs K3,
expectation runs the fused K2 kernel and get_ground_state / get_gap run the
device-resident Lanczos loop on top of K1.
"""
import functools
import os
import itertools

import numpy
import scipy.sparse
import scipy.sparse.linalg

from openfermion_b200 import _lib, ops, sectors
from openfermion_b200.compiler import compile_fermion_terms, compile_qubit_operator
from openfermion_b200.linalg import krylov
from openfermion_b200.linalg.linear_qubit_operator import (LinearQubitOperator,
                                                           SectorLinearQubitOperator)
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


def jordan_wigner_ladder_sparse(n_qubits, tensor_factor, ladder_type):
    """CSC matrix of one ladder operator, a_j^dagger (ladder_type truthy) or a_j on mode
    `tensor_factor` (sparse_tools.py:49-73: Z_0..Z_{j-1} (X_j -/+ iY_j)/2).  The reference
    chains Kronecker products of 2x2 blocks; here it is the one-factor case of the fermion
    route, built by the CSC kernels (K4)."""
    if not 0 <= tensor_factor < n_qubits:
        raise ValueError('tensor_factor must address one of the n_qubits modes.')
    operator = ops.FermionOperator(((int(tensor_factor), 1 if ladder_type else 0),), 1.0)
    return jordan_wigner_sparse(operator, n_qubits)


def wrapped_kronecker(operator_1, operator_2):
    """Kronecker product of two sparse matrices as CSC (sparse_tools.py:39-41).  A generic
    scipy helper of the reference, kept for callers that assemble their own operators; the
    Hamiltonian builders above never form Kronecker chains."""
    return scipy.sparse.kron(operator_1, operator_2, 'csc')


def kronecker_operators(*args):
    """Kronecker product of a list of sparse matrices (sparse_tools.py:44-46); like the
    reference it takes the list as its single positional argument."""
    return functools.reduce(wrapped_kronecker, *args)


def _embed_two_level_block(entries, first_qubit, n_block_qubits, n_qubits):
    """CSC of  1 (x) B (x) 1  for a small block B given by its stored entries (row, col, value),
    acting on qubits first_qubit .. first_qubit + n_block_qubits - 1: every stored entry of B
    -- explicit zeros included, as scipy's Kronecker product keeps them -- appears once per
    value of the other qubits.  Index arithmetic only; no Kronecker chain."""
    right_bits = n_qubits - first_qubit - n_block_qubits
    left = numpy.arange(1 << first_qubit, dtype=numpy.int64)[:, None, None]
    right = numpy.arange(1 << right_bits, dtype=numpy.int64)[None, None, :]
    block_rows = numpy.array([e[0] for e in entries], dtype=numpy.int64)[None, :, None]
    block_cols = numpy.array([e[1] for e in entries], dtype=numpy.int64)[None, :, None]
    values = numpy.array([e[2] for e in entries], dtype=numpy.complex128)[None, :, None]
    rows = (((left << n_block_qubits) | block_rows) << right_bits) | right
    cols = (((left << n_block_qubits) | block_cols) << right_bits) | right
    data = numpy.broadcast_to(values, rows.shape)
    dim = 1 << n_qubits
    return scipy.sparse.csc_matrix((data.ravel(), (rows.ravel(), cols.ravel())), shape=(dim, dim),
                                   dtype=numpy.complex128)


def jw_sparse_givens_rotation(i, j, theta, phi, n_qubits):
    """Matrix on the full 2^n space of a Givens rotation of the adjacent modes i, j = i + 1 in
    the Jordan-Wigner encoding (sparse_tools.py:516-545).  A gate-building helper of the
    reference (host, scipy there as here); same stored entries, explicit zeros included."""
    if j != i + 1:
        raise ValueError('Only adjacent modes can be rotated.')
    if j > n_qubits - 1:
        raise ValueError('Too few qubits requested.')
    cosine, sine, phase = numpy.cos(theta), numpy.sin(theta), numpy.exp(1.0j * phi)
    block = [(0, 0, 1.0), (1, 1, phase * cosine), (1, 2, -phase * sine), (2, 1, sine), (2, 2, cosine),
             (3, 3, phase)]
    return _embed_two_level_block(block, i, 2, n_qubits)


def jw_sparse_particle_hole_transformation_last_mode(n_qubits):
    """Particle-hole transformation of the last mode: 1 (x) X (sparse_tools.py:548-554)."""
    return _embed_two_level_block([(0, 1, 1.0), (1, 0, 1.0)], n_qubits - 1, 1, n_qubits)


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


def _tensor_rows(operator):
    """(n_qubits, projected rows) of a PolynomialTensor-like operator whose tensors are a constant,
    a one-body and a two-body part (compiler.compile_interaction_tensors), else None."""
    tensors = getattr(operator, 'n_body_tensors', None)
    if not isinstance(tensors, dict) or os.environ.get('OFB_TENSOR_ROWS', '1') == '0':
        return None
    from openfermion_b200.compiler import compile_interaction_tensors
    return compile_interaction_tensors(tensors, None, ops.EQ_TOLERANCE)


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
        rows = _tensor_rows(operator)
        if rows is not None:
            # InteractionOperator-like tensors: the projected rows are built array-wise, identical
            # to the rows of the FermionOperator the reference converts to first (a Python loop
            # over N^4 index tuples: 10 of the 14 ms of this call for LiH, seconds at 20 qubits)
            plan = PauliPlan.projected_rows(rows[0], *rows[1])
            try:
                return plan.csc_host()
            finally:
                plan.close()
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
        if isinstance(operator, SectorLinearQubitOperator):
            if state.size != operator.shape[1]:
                raise ValueError('Dimension mismatch')
            return operator.sector.expectation_host(operator.apply_plan, state)
        applied = krylov.apply_any(operator, state)
        if len(state.shape) == 1:
            return numpy.dot(numpy.conjugate(state), applied)
        return numpy.dot(numpy.conjugate(state.T), applied)[0, 0]
    raise ValueError('Input state must be a numpy array or a sparse matrix.')


def variance(operator, state):
    """<H^2> - <H>^2 (sparse_tools.py:674-690).  For the matrix-free operators the reference's
    `operator**2` is scipy's power operator, i.e. two host matvecs plus a third for <H>; here the
    state is uploaded once and w = H v, u = H w, <v|u> - <v|w>^2 are formed on the device."""
    if (isinstance(state, numpy.ndarray) and state.ndim == 1
            and isinstance(operator, (LinearQubitOperator, SectorLinearQubitOperator))):
        if state.size != operator.shape[1]:
            raise ValueError('Dimension mismatch')
        dev_op = krylov.as_device_operator(operator)
        n = dev_op.dim
        v = _lib.to_device(numpy.ascontiguousarray(state, dtype=numpy.complex128))
        w = _lib.DeviceBuffer(n * krylov.C128)
        u = _lib.DeviceBuffer(n * krylov.C128)
        try:
            dev_op.apply(v.ptr, w.ptr)
            dev_op.apply(w.ptr, u.ptr)
            return krylov.dotc(n, v.ptr, u.ptr) - krylov.dotc(n, v.ptr, w.ptr) ** 2
        finally:
            for buf in (v, w, u):
                buf.free()
    return expectation(operator**2, state) - expectation(operator, state) ** 2


def expectation_computational_basis_state(operator, computational_basis_state):
    """<state|operator|state> for a normal-ordered FermionOperator and one computational basis
    state (sparse_tools.py:693-737).  Host arithmetic on the symbolic terms, as in the
    reference: only number-like terms p^ p and q^ p^ q p of occupied orbitals contribute.  The
    state is a list of occupations or a sparse / dense vector with a single non-zero entry,
    whose index is read least-significant bit = orbital 0 (the reference's convention here)."""
    if ops.is_qubit_operator(operator):
        raise NotImplementedError('Not yet implemented for QubitOperators.')
    if not ops.is_fermion_operator(operator):
        raise TypeError('operator must be a FermionOperator.')
    checker = getattr(operator, 'is_normal_ordered', None)
    if not (checker() if callable(checker) else ops.is_normal_ordered(operator)):
        raise ValueError('operator must be a normal ordered.')
    occupied = computational_basis_state
    if not isinstance(occupied, list):
        index = int(occupied.nonzero()[0][0])
        occupied = [(index >> orbital) & 1 == 1 for orbital in range(index.bit_length())]
    terms = operator.terms
    value = terms.get((), 0.0)
    filled = [orbital for orbital, flag in enumerate(occupied) if flag]
    for position, p in enumerate(filled):
        value += terms.get(((p, 1), (p, 0)), 0.0)
        for q in filled[position + 1:]:
            value -= terms.get(((q, 1), (p, 1), (q, 0), (p, 0)), 0.0)
    return value


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


def jw_configuration_state(occupied_orbitals, n_qubits):
    """Basis state with the given orbitals occupied, float64[2^n] (sparse_tools.py:250-264)."""
    one_index = sum(2 ** (n_qubits - 1 - i) for i in occupied_orbitals)
    basis_vector = numpy.zeros(2**n_qubits, dtype=float)
    basis_vector[one_index] = 1
    return basis_vector


def jw_hartree_fock_state(n_electrons, n_orbitals):
    """Hartree-Fock determinant in the JW encoding (sparse_tools.py:267-270)."""
    return jw_configuration_state(range(n_electrons), n_orbitals)


def get_density_matrix(states, probabilities):
    """sum_k p_k |psi_k><psi_k| as scipy CSC (sparse_tools.py:557-563)."""
    n_qubits = states[0].shape[0]
    density_matrix = scipy.sparse.csc_matrix((n_qubits, n_qubits), dtype=complex)
    for state, probability in zip(states, probabilities):
        state = scipy.sparse.csc_matrix(state.reshape((len(state), 1)))
        density_matrix = density_matrix + probability * state * state.getH()
    return density_matrix


def _is_hermitian_sparse(operator):
    difference = operator - operator.getH()
    return difference.nnz == 0 or numpy.max(numpy.abs(difference.data)) <= ops.EQ_TOLERANCE


def sparse_eigenspectrum(sparse_operator):
    """Dense diagonalisation of a materialised operator, sorted (sparse_tools.py:615-626)."""
    dense_operator = sparse_operator.toarray()
    if _is_hermitian_sparse(scipy.sparse.csc_matrix(sparse_operator)):
        spectrum = numpy.linalg.eigvalsh(dense_operator)
    else:
        spectrum = numpy.linalg.eigvals(dense_operator)
    return numpy.sort(spectrum)


def eigenspectrum(operator, n_qubits=None):
    """Full spectrum of a symbolic operator: device CSC build, then the reference's dense
    solve (sparse_tools.py:590-612; cubic in 2^n, small systems only)."""
    names = {c.__name__ for c in type(operator).__mro__}
    if names & {'BosonOperator', 'QuadOperator'}:
        raise TypeError('Operator of invalid type.')
    return sparse_eigenspectrum(get_sparse_operator(operator, n_qubits))


def inner_product(state_1, state_2):
    """conj(state_1) . state_2 (sparse_tools.py:1098-1100)."""
    return numpy.dot(state_1.conjugate(), state_2)


# ----------------------------------------------------------------------------
# Jordan-Wigner symmetry sectors
# (sparse_tools.py:273-513 and :1282-1597).  The reference materialises the 2^n x 2^n matrix
# and slices it; that route is kept for already materialised inputs.  Symbolic operators
# (QubitOperator / FermionOperator / LinearQubitOperator) never leave the sector: the index
# map lives on the device (openfermion_b200/sectors.py, csrc/sector.cu) and the restricted
# operator is either applied matrix-free or assembled directly as a dim x dim CSC.
# ----------------------------------------------------------------------------
def jw_number_indices(n_electrons, n_qubits):
    """Indices of all basis states with n_electrons occupied modes, in the reference's order
    (itertools.combinations over bit positions, sparse_tools.py:273-292)."""
    return sectors.number_indices_list(n_electrons, n_qubits)


def jw_sz_indices(sz_value, n_qubits, n_electrons=None, up_index=None, down_index=None):
    """Indices of basis states with fixed S_z (sparse_tools.py:295-372); up = 2i, down = 2i+1
    by default (utils/indexing.py:17,29)."""
    return sectors.sz_indices_list(sz_value, n_qubits, n_electrons, up_index, down_index)


def _is_symbolic(operator):
    return (ops.is_qubit_operator(operator) or ops.is_fermion_operator(operator)
            or isinstance(operator, LinearQubitOperator))


def _sector_plan(operator, n_qubits):
    """Compiled plan of a symbolic operator for the sector kernels: (plan, owned)."""
    if isinstance(operator, LinearQubitOperator):
        if n_qubits is not None and n_qubits != operator.n_qubits:
            raise ValueError('Invalid number of qubits specified.')
        return operator.plan, False
    needed = count_qubits(operator)
    if n_qubits is None:
        n_qubits = needed
    if n_qubits < needed:
        raise ValueError('Invalid number of qubits specified.')
    if ops.is_fermion_operator(operator):
        return PauliPlan.projected_rows(n_qubits, *compile_fermion_terms(operator.terms, n_qubits)), True
    x, z, c = compile_qubit_operator(operator, n_qubits)
    return PauliPlan(n_qubits, x, z, c), True


def _restricted_csc(operator, n_qubits, make_sector):
    plan, owned = _sector_plan(operator, n_qubits)
    sector = make_sector(plan.n_qubits)
    try:
        return sector.csc_host(plan)
    finally:
        sector.close()
        if owned:
            plan.close()


def jw_number_restrict_operator(operator, n_electrons, n_qubits=None):
    """Operator restricted to n_electrons particles (sparse_tools.py:375-392).  Matrices are
    sliced like the reference does; symbolic operators are assembled inside the sector on the
    device (the full matrix is never built)."""
    if _is_symbolic(operator):
        return _restricted_csc(operator, n_qubits,
                               lambda n: sectors.Sector.number(n, n_electrons))
    if n_qubits is None:
        n_qubits = int(numpy.log2(operator.shape[0]))
    select = jw_number_indices(n_electrons, n_qubits)
    return operator[numpy.ix_(select, select)]


def jw_sz_restrict_operator(operator, sz_value, n_electrons=None, n_qubits=None, up_index=None,
                            down_index=None):
    """Operator restricted to a given S_z (and optionally particle number)
    (sparse_tools.py:395-423)."""
    if _is_symbolic(operator):
        return _restricted_csc(operator, n_qubits,
                               lambda n: sectors.Sector.sz(n, sz_value, n_electrons, up_index,
                                                           down_index))
    if n_qubits is None:
        n_qubits = int(numpy.log2(operator.shape[0]))
    select = jw_sz_indices(sz_value, n_qubits, n_electrons=n_electrons, up_index=up_index,
                           down_index=down_index)
    return operator[numpy.ix_(select, select)]


def jw_number_restrict_state(state, n_electrons, n_qubits=None):
    """state[jw_number_indices] (sparse_tools.py:426-442)."""
    if n_qubits is None:
        n_qubits = int(numpy.log2(state.shape[0]))
    return state[jw_number_indices(n_electrons, n_qubits)]


def jw_sz_restrict_state(state, sz_value, n_electrons=None, n_qubits=None, up_index=None,
                         down_index=None):
    """state[jw_sz_indices] (sparse_tools.py:445-476)."""
    if n_qubits is None:
        n_qubits = int(numpy.log2(state.shape[0]))
    return state[jw_sz_indices(sz_value, n_qubits, n_electrons=n_electrons, up_index=up_index,
                               down_index=down_index)]


def jw_get_ground_state_at_particle_number(sparse_operator, particle_number, sz_value=None,
                                           expand=True):
    """Lowest eigenpair inside a particle-number sector, expanded back to 2^n amplitudes
    (sparse_tools.py:477-513).

    Beyond the reference: a symbolic operator or LinearQubitOperator is solved matrix-free
    inside the sector (device Lanczos on the sector kernel); `sz_value` narrows the sector
    to fixed S_z as well, and expand=False returns the dim-long sector vector (the 2^n
    expansion of a 32-qubit state is 64 GiB)."""
    if ops.is_fermion_operator(sparse_operator):
        # the matrix-free sector kernel runs on Pauli tables; a FermionOperator is assembled
        # inside the sector from its projected rows (as jw_number_restrict_operator does) and
        # solved as a sparse matrix
        n_qubits = count_qubits(sparse_operator)
        if sz_value is None:
            restricted = jw_number_restrict_operator(sparse_operator, particle_number, n_qubits)
            select = jw_number_indices(particle_number, n_qubits)
        else:
            restricted = jw_sz_restrict_operator(sparse_operator, sz_value, particle_number, n_qubits)
            select = jw_sz_indices(sz_value, n_qubits, n_electrons=particle_number)
        if restricted.shape[0] - 1 <= 1:
            eigvals, eigvecs = numpy.linalg.eigh(restricted.toarray())
            energy, state = eigvals[0], eigvecs[:, 0]
        else:
            energy, state = get_ground_state(scipy.sparse.csc_matrix(restricted))
        if not expand:
            return energy, state
        expanded = numpy.zeros(2**n_qubits, dtype=complex)
        expanded[select] = state
        return energy, expanded
    if _is_symbolic(sparse_operator):
        from openfermion_b200.linalg.linear_qubit_operator import SectorLinearQubitOperator
        restricted = SectorLinearQubitOperator(sparse_operator, n_electrons=particle_number,
                                               sz_value=sz_value)
        if restricted.shape[0] - 1 <= 1:
            dense = restricted.matmat(numpy.eye(restricted.shape[0], dtype=complex))
            eigvals, eigvecs = numpy.linalg.eigh(dense)
            energy, state = eigvals[0], eigvecs[:, 0]
        else:
            energy, state = get_ground_state(restricted)
        if not expand:
            return energy, state
        return energy, restricted.sector.expand_state_host(state)
    if sz_value is not None:
        raise ValueError('sz_value needs a symbolic operator or a LinearQubitOperator.')
    n_qubits = int(numpy.log2(sparse_operator.shape[0]))
    restricted = jw_number_restrict_operator(sparse_operator, particle_number, n_qubits)
    if restricted.shape[0] - 1 <= 1:
        eigvals, eigvecs = numpy.linalg.eigh(restricted.toarray())
        energy, state = eigvals[0], eigvecs[:, 0]
    else:
        energy, state = get_ground_state(scipy.sparse.csc_matrix(restricted))
    if not expand:
        return energy, state
    expanded = numpy.zeros(2**n_qubits, dtype=complex)
    expanded[jw_number_indices(particle_number, n_qubits)] = state
    return energy, expanded


def _excited_determinants(reference, occupied, unoccupied, order, n_qubits):
    for occ in itertools.combinations(occupied, order):
        cleared = reference
        for mode in occ:
            cleared ^= 1 << (n_qubits - 1 - mode)
        for unocc in itertools.combinations(unoccupied, order):
            state = cleared
            for mode in unocc:
                state ^= 1 << (n_qubits - 1 - mode)
            yield state


def _number_preserving_basis(reference_determinant, excitation_level, spin_preserving):
    """Integer encodings (mode i = bit n-1-i) of the determinants in the order of
    _iterate_basis_ (sparse_tools.py:1370-1475): by excitation rank from the reference and,
    with spin_preserving, by (alpha rank, beta rank) inside each rank."""
    reference_determinant = [bool(b) for b in reference_determinant]
    n = len(reference_determinant)
    ref_int = sum(1 << (n - 1 - i) for i, b in enumerate(reference_determinant) if b)
    basis = []
    if not spin_preserving:
        occupied = [i for i, b in enumerate(reference_determinant) if b]
        unoccupied = [i for i, b in enumerate(reference_determinant) if not b]
        for order in range(excitation_level + 1):
            basis.extend(_excited_determinants(ref_int, occupied, unoccupied, order, n))
        return basis
    occ_a = [i for i in range(0, n, 2) if reference_determinant[i]]
    unocc_a = [i for i in range(0, n, 2) if not reference_determinant[i]]
    occ_b = [i for i in range(1, n, 2) if reference_determinant[i]]
    unocc_b = [i for i in range(1, n, 2) if not reference_determinant[i]]
    alpha_level = min(len(occ_a), excitation_level)
    beta_level = min(len(occ_b), excitation_level)
    for order in range(excitation_level + 1):
        for alpha_order in range(alpha_level + 1):
            beta_order = order - alpha_order
            if beta_order < 0 or beta_order > beta_level:
                continue
            # the alpha excitation is the outer loop, the beta excitation the inner one
            beta_flips = [s ^ ref_int for s in
                          _excited_determinants(ref_int, occ_b, unocc_b, beta_order, n)]
            for alpha_state in _excited_determinants(ref_int, occ_a, unocc_a, alpha_order, n):
                basis.extend(alpha_state ^ flip for flip in beta_flips)
    return basis


def get_number_preserving_sparse_operator(fermion_op, num_qubits, num_electrons,
                                          spin_preserving=False, reference_determinant=None,
                                          excitation_level=None):
    """CSC of a FermionOperator inside a particle-number (and optionally S_z, optionally
    excitation-truncated) sector, basis ordered by excitation rank from the reference
    determinant (sparse_tools.py:1282-1369).  The reference loops over terms and basis states
    in Python (:1478-1597); here the normal-ordered terms become projected scatter rows and
    the sector CSC kernels place them through a device-side search of the basis."""
    if reference_determinant is None:
        reference_determinant = [i < num_electrons for i in range(num_qubits)]
    reference_determinant = list(numpy.asarray(reference_determinant).astype(bool))
    if excitation_level is None:
        excitation_level = num_electrons
    basis = _number_preserving_basis(reference_determinant, excitation_level, spin_preserving)
    converted = ops.FermionOperator()
    converted.terms = dict(fermion_op.terms)
    ordered = ops.normal_ordered(converted)
    is_complex = False
    for term, coefficient in ordered.terms.items():
        if sum(1 if kind else -1 for _, kind in term) != 0:
            raise ValueError("The supplied operator doesn't preserve particle number")
        if isinstance(coefficient, (complex, numpy.complexfloating)):
            is_complex = True
    plan = PauliPlan.projected_rows(num_qubits, *compile_fermion_terms(ordered.terms, num_qubits))
    sector = sectors.Sector.from_list(num_qubits, basis)
    try:
        return sector.csc_host(plan, dtype=numpy.complex128 if is_complex else numpy.float64)
    finally:
        sector.close()
        plan.close()
