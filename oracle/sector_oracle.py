"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY: symmetry-sector functions of the reference.

Restates, with numpy/scipy on the CPU,

    /root/reference/src/openfermion/linalg/sparse_tools.py
        :273-292    jw_number_indices
        :295-372    jw_sz_indices
        :375-476    jw_number_restrict_operator / jw_sz_restrict_operator / *_restrict_state
        :477-513    jw_get_ground_state_at_particle_number
        :1282-1597  get_number_preserving_sparse_operator and its helpers

so the sector kernels (csrc/sector.cu) can be parity-checked on a box without the reference.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
Parity pinning: oracle/make_golden.py asserts every function below against the real
reference (imported through oracle/ref_shim.py) and commits the reference's outputs to
tests/golden/sector_cases.npz; tests/test_oracle_golden.py re-checks the oracle against them.
"""
import itertools

import numpy
import scipy.sparse
import scipy.sparse.linalg


def jw_number_indices(n_electrons, n_qubits):
    """sparse_tools.py:273-292: combinations of BIT positions, index = sum 2^k."""
    return [sum(2**k for k in occupation)
            for occupation in itertools.combinations(range(n_qubits), n_electrons)]


def jw_sz_indices(sz_value, n_qubits, n_electrons=None, up_index=None, down_index=None):
    """sparse_tools.py:295-372: spin-up site i is qubit up_index(i) (default 2i), spin-down
    2i+1; qubit q is index bit n_qubits-1-q; the first species' occupations are the outer
    loop."""
    up_index = up_index or (lambda i: 2 * i)
    down_index = down_index or (lambda i: 2 * i + 1)
    if n_qubits % 2 != 0:
        raise ValueError('Number of qubits must be even')
    if not (2.0 * sz_value).is_integer():
        raise ValueError('Sz value must be an integer or half-integer')
    n_sites = n_qubits // 2
    sz_integer = int(2.0 * sz_value)
    weight = lambda qubits: sum(2 ** (n_qubits - 1 - q) for q in qubits)  # noqa: E731
    indices = []
    if n_electrons is not None:
        if (n_electrons + sz_integer) % 2 != 0 or n_electrons < abs(sz_integer):
            raise ValueError('The specified particle number and sz value are incompatible.')
        num_up = (n_electrons + sz_integer) // 2
        plan = [(up_index, num_up, down_index, n_electrons - num_up)]
    else:
        first, second = (down_index, up_index) if sz_integer < 0 else (up_index, down_index)
        plan = [(first, n, second, n - abs(sz_integer)) for n in range(abs(sz_integer), n_sites + 1)]
    for first, n_first, second, n_second in plan:
        seconds = [weight([second(i) for i in occ])
                   for occ in itertools.combinations(range(n_sites), n_second)]
        for occ in itertools.combinations(range(n_sites), n_first):
            w = weight([first(i) for i in occ])
            indices.extend(w + s for s in seconds)
    return indices


def restrict_operator(operator, indices):
    """operator[numpy.ix_(indices, indices)] (sparse_tools.py:391-392, 420-423)."""
    return operator[numpy.ix_(indices, indices)]


def ground_state_at_particle_number(sparse_operator, particle_number):
    """sparse_tools.py:477-513."""
    n_qubits = int(numpy.log2(sparse_operator.shape[0]))
    indices = jw_number_indices(particle_number, n_qubits)
    restricted = restrict_operator(sparse_operator, indices)
    if restricted.shape[0] - 1 <= 1:
        values, vectors = numpy.linalg.eigh(restricted.toarray())
    else:
        values, vectors = scipy.sparse.linalg.eigsh(restricted, k=1, which='SA')
    expanded = numpy.zeros(2**n_qubits, dtype=complex)
    expanded[indices] = vectors[:, 0]
    return values[0], expanded


# ---- get_number_preserving_sparse_operator ---------------------------------------------
def iterate_basis(reference_determinant, excitation_level, spin_preserving):
    """Determinants (bool arrays) in the order of _iterate_basis_ (sparse_tools.py:1370-1475)."""
    reference_determinant = numpy.asarray(reference_determinant, dtype=bool)

    def excite(pools):
        # pools: [(occupied modes, unoccupied modes, order)], first pool = outermost loops
        loops = []
        for occupied, unoccupied, order in pools:
            loops.append(itertools.combinations(occupied, order))
            loops.append(itertools.combinations(unoccupied, order))
        for choice in itertools.product(*loops):
            state = reference_determinant.copy()
            for k, modes in enumerate(choice):
                state[list(modes)] = bool(k % 2)  # even slots empty occupied modes, odd slots fill
            yield state

    if not spin_preserving:
        occupied = numpy.where(reference_determinant)[0]
        unoccupied = numpy.where(~reference_determinant)[0]
        for order in range(excitation_level + 1):
            yield from excite([(occupied, unoccupied, order)])
        return
    occ_a = numpy.where(reference_determinant[::2])[0] * 2
    unocc_a = numpy.where(~reference_determinant[::2])[0] * 2
    occ_b = numpy.where(reference_determinant[1::2])[0] * 2 + 1
    unocc_b = numpy.where(~reference_determinant[1::2])[0] * 2 + 1
    max_a = min(int(numpy.sum(reference_determinant[::2])), excitation_level)
    max_b = min(int(numpy.sum(reference_determinant[1::2])), excitation_level)
    for order in range(excitation_level + 1):
        for order_a in range(max_a + 1):
            order_b = order - order_a
            if order_b < 0 or order_b > max_b:
                continue
            yield from excite([(occ_a, unocc_a, order_a), (occ_b, unocc_b, order_b)])


def build_term_op(term, state_array, int_state_array, sorting_indices):
    """Sparse matrix of one normal-ordered ladder product in the given determinant basis
    (sparse_tools.py:1478-1597)."""
    space_size = state_array.shape[0]
    must_be_filled, must_be_empty = [], []
    delta = 0
    for mode, kind in reversed(term):
        if kind == 0:
            must_be_filled.append(mode)
            delta -= 1
        else:
            if mode not in must_be_filled:
                must_be_empty.append(mode)
            delta += 1
    if delta != 0:
        raise ValueError("The supplied operator doesn't preserve particle number")
    candidates = numpy.where(numpy.all(state_array[:, must_be_filled], axis=1)
                             & ~numpy.any(state_array[:, must_be_empty], axis=1))[0]
    weights = 1 << numpy.arange(state_array.shape[1])[::-1]
    data, rows, cols = [], [], []
    for col in candidates:
        target = state_array[col].copy()
        sign = 1
        for mode, _ in reversed(term):
            sign *= (-1) ** int(numpy.sum(target[:mode]))
            target[mode] = not target[mode]
        code = target.dot(weights)
        where = numpy.searchsorted(int_state_array, code, sorter=sorting_indices)
        row = sorting_indices[min(where, space_size - 1)]
        if where < space_size and int_state_array[row] == code:
            data.append(sign)
            rows.append(row)
            cols.append(col)
    return scipy.sparse.csc_matrix((numpy.asarray(data), (numpy.asarray(rows), numpy.asarray(cols))),
                                   shape=(space_size, space_size))


def get_number_preserving_sparse_operator(normal_ordered_terms, num_qubits, num_electrons,
                                          spin_preserving=False, reference_determinant=None,
                                          excitation_level=None):
    """sparse_tools.py:1282-1369 on an ALREADY normal-ordered `terms` dict (the reference
    normal-orders first, :1352; openfermion_b200.ops.normal_ordered reproduces that step
    bit for bit, see make_golden.py)."""
    if reference_determinant is None:
        reference_determinant = numpy.array([i < num_electrons for i in range(num_qubits)])
    else:
        reference_determinant = numpy.asarray(reference_determinant)
    if excitation_level is None:
        excitation_level = num_electrons
    state_array = numpy.asarray(list(iterate_basis(reference_determinant, excitation_level, spin_preserving)))
    int_state_array = state_array.dot(1 << numpy.arange(state_array.shape[1])[::-1])
    sorting_indices = numpy.argsort(int_state_array)
    space_size = state_array.shape[0]
    total = scipy.sparse.csc_matrix((space_size, space_size), dtype=float)
    for term, coefficient in normal_ordered_terms.items():
        if len(term) == 0:
            total += coefficient * scipy.sparse.identity(space_size, dtype=float, format='csc')
        else:
            total += coefficient * build_term_op(term, state_array, int_state_array, sorting_indices)
    return total
