"""Term compiler: symbolic `terms` dicts -> packed mask tables for libofb200.

Qubit q is index bit (n_qubits - 1 - q), as in the reference
(linalg/linear_qubit_operator.py:78-84,126).
"""
import numbers

import numpy

from openfermion_b200.ops import count_qubits


def _as_complex(coefficient):
    if isinstance(coefficient, numbers.Number):
        return complex(coefficient)
    try:  # numpy scalars / 0-d arrays
        return complex(numpy.asarray(coefficient).item())
    except Exception:
        raise TypeError(
            'Only numeric coefficients can be compiled for the numeric path, got {!r} '
            '(symbolic coefficients are not supported).'.format(type(coefficient).__name__)
        )


def compile_qubit_terms(terms, n_qubits):
    """QubitOperator.terms -> (x_mask u64[T], z_mask u64[T], coeff c128[T]) in dict order.

    Mask folding is the reference's (linear_qubit_operator.py:122-134): X and Y set the
    x bit, Y and Z set the z bit.  The i^y phase is applied inside the library.
    """
    count = len(terms)
    x = numpy.zeros(count, dtype=numpy.uint64)
    z = numpy.zeros(count, dtype=numpy.uint64)
    c = numpy.zeros(count, dtype=numpy.complex128)
    top = n_qubits - 1
    for k, (term, coefficient) in enumerate(terms.items()):
        xm = zm = 0
        for qubit, action in term:
            if qubit > top:
                raise ValueError('Invalid number of qubits specified.')
            bit = 1 << (top - qubit)
            if action == 'X':
                xm ^= bit
            elif action == 'Y':
                xm ^= bit
                zm ^= bit
            elif action == 'Z':
                zm ^= bit
            else:
                raise ValueError('Invalid Pauli action {!r}'.format(action))
        x[k] = xm
        z[k] = zm
        c[k] = _as_complex(coefficient)
    return x, z, c


def compile_qubit_operator(operator, n_qubits):
    """compile_qubit_terms for any QubitOperator-like object; array-form operators
    (hamiltonians.MaskOperator) skip the dict walk."""
    if hasattr(operator, 'index_masks'):
        x, z = operator.index_masks(n_qubits)
        return x, z, numpy.asarray(operator.coeff, dtype=numpy.complex128)
    return compile_qubit_terms(operator.terms, n_qubits)


def compile_fermion_terms(terms, n_qubits):
    """FermionOperator.terms -> projected scatter rows for the Jordan-Wigner matrix.

    Derived from the reference's ladder matrices (linalg/sparse_tools.py:49-73: a_p =
    Z_0..Z_{p-1} (X_p + iY_p)/2 lowers index bit n-1-p with sign (-1)^{occupied modes
    before p}) and the right-to-left action of a product on a basis ket
    (sparse_tools.py:110-116).  A term contributes
        H[c ^ x, c] += coeff * (-1)^popc(c & z)   for columns with (c & occ_mask) == occ_val
    or nothing at all when it hits an occupied mode with a raising operator (or an empty
    one with a lowering operator) regardless of c.  Zero coefficients are skipped like
    the reference's `if coefficient:` (sparse_tools.py:118).
    Returns (occ_mask, occ_val, x_mask, z_mask, coeff) arrays in dict order.
    """
    occ_m, occ_v, xs, zs, cs = [], [], [], [], []
    top = n_qubits - 1
    full = (1 << n_qubits) - 1
    for term, coefficient in terms.items():
        if not coefficient:
            continue
        flip = 0       # bits flipped so far (state = column ^ flip)
        z_mask = 0     # sign bits that depend on the column
        sign = 0       # constant sign from bits already flipped
        need_mask = 0  # occupation constraints on the column
        need_val = 0
        dead = False
        for mode, kind in reversed(term):
            if mode > top:
                raise ValueError('Invalid number of qubits specified.')
            bit = 1 << (top - mode)
            above = full ^ ((bit << 1) - 1)  # index bits of modes before `mode`
            # current occupation of this mode is column_bit ^ flip_bit; a raising
            # operator needs it empty, a lowering operator needs it filled
            want_now = 0 if kind else 1
            want_col = want_now ^ (1 if flip & bit else 0)
            if need_mask & bit:
                if (1 if need_val & bit else 0) != want_col:
                    dead = True
                    break
            else:
                need_mask |= bit
                if want_col:
                    need_val |= bit
            z_mask ^= above
            sign ^= bin(flip & above).count('1') & 1
            flip ^= bit
        if dead:
            continue
        c = _as_complex(coefficient)
        occ_m.append(need_mask)
        occ_v.append(need_val)
        xs.append(flip)
        zs.append(z_mask)
        cs.append(-c if sign else c)
    return (numpy.array(occ_m, dtype=numpy.uint64), numpy.array(occ_v, dtype=numpy.uint64),
            numpy.array(xs, dtype=numpy.uint64), numpy.array(zs, dtype=numpy.uint64),
            numpy.array(cs, dtype=numpy.complex128))


def _parity_u64(values):
    """parity of the set bits of a uint64 array (portable: no numpy.bitwise_count)."""
    v = values.copy()
    for shift in (32, 16, 8, 4, 2, 1):
        v ^= v >> numpy.uint64(shift)
    return (v & numpy.uint64(1)).astype(bool)


def compile_ladder_products(modes, kinds, coefficients, n_qubits):
    """compile_fermion_terms for MANY products with the same pattern of raising / lowering
    operators at once: modes[N, K] (mode of factor k of product i), kinds[K] (1 raising,
    0 lowering), coefficients[N].  Same bit logic, same skip rules (zero coefficients, products
    that vanish on every column), same output order as a loop over the N products."""
    modes = numpy.asarray(modes, dtype=numpy.int64).reshape(len(coefficients), -1)
    coeff = numpy.asarray(coefficients, dtype=numpy.complex128)
    if modes.size and int(modes.max()) > n_qubits - 1:
        raise ValueError('Invalid number of qubits specified.')
    count = len(coeff)
    top = numpy.uint64(max(n_qubits - 1, 0))  # unused when there are no factors (a constant)
    full = numpy.uint64((1 << n_qubits) - 1)
    one = numpy.uint64(1)
    zero = numpy.zeros(count, dtype=numpy.uint64)
    flip, z_mask, need_mask, need_val = zero.copy(), zero.copy(), zero.copy(), zero.copy()
    sign = numpy.zeros(count, dtype=bool)
    dead = numpy.zeros(count, dtype=bool)
    for k in reversed(range(modes.shape[1])):
        bit = one << (top - modes[:, k].astype(numpy.uint64))
        above = full ^ ((bit << one) - one)
        want_col = numpy.full(count, not kinds[k]) ^ ((flip & bit) != 0)
        has = (need_mask & bit) != 0
        dead |= has & (((need_val & bit) != 0) != want_col)
        fresh = ~has
        need_mask |= numpy.where(fresh, bit, zero)
        need_val |= numpy.where(fresh & want_col, bit, zero)
        z_mask ^= above
        sign ^= _parity_u64(flip & above)
        flip ^= bit
    keep = ~dead & (coeff != 0)
    signed = numpy.where(sign, -coeff, coeff)
    return need_mask[keep], need_val[keep], flip[keep], z_mask[keep], signed[keep]


def compile_interaction_tensors(n_body_tensors, n_qubits, tolerance):
    """Projected rows of constant + sum_pq h_pq p^ q + sum_pqrs g_pqrs p^ q^ r s straight from
    the tensors, identical to compile_fermion_terms(get_fermion_operator(tensors).terms): the
    reference's conversion (transforms/opconversions/conversions.py:117-121) visits the constant,
    then each tensor's non-zero entries in index-product order and drops what `+=` drops
    (|coefficient| < EQ_TOLERANCE).  n_qubits None: 1 + the highest mode of a surviving term, as
    count_qubits of that FermionOperator.  Returns (n_qubits, rows) or None for tensors of
    another shape."""
    keys = set(n_body_tensors.keys())
    if not keys <= {(), (1, 0), (1, 1, 0, 0)}:
        return None
    blocks = []  # (modes[N, K], kinds, coefficients[N])
    constant = n_body_tensors.get((), 0.0)
    if constant and abs(constant) >= tolerance:
        blocks.append((numpy.zeros((1, 0), dtype=numpy.int64), (), numpy.array([constant])))
    for key in ((1, 0), (1, 1, 0, 0)):
        if key not in keys:
            continue
        tensor = numpy.asarray(n_body_tensors[key])
        index = numpy.nonzero(tensor)  # C order = itertools.product order
        values = tensor[index]
        big = numpy.abs(values) >= tolerance
        if numpy.any(big):
            blocks.append((numpy.stack([axis[big] for axis in index], axis=1), key, values[big]))
    needed = max([int(modes.max()) + 1 for modes, _, _ in blocks if modes.size], default=0)
    if n_qubits is None:
        n_qubits = needed
    elif needed > n_qubits:
        raise ValueError('Invalid number of qubits specified.')
    parts = [compile_ladder_products(modes, kinds, values, n_qubits) for modes, kinds, values in blocks]
    if not parts:
        empty = numpy.zeros(0, dtype=numpy.uint64)
        return n_qubits, (empty, empty.copy(), empty.copy(), empty.copy(), numpy.zeros(0, dtype=numpy.complex128))
    return n_qubits, tuple(numpy.concatenate([part[k] for part in parts]) for k in range(5))


def resolve_n_qubits(operator, n_qubits, message='Invalid number of qubits specified.'):
    needed = count_qubits(operator)
    if n_qubits is None:
        return needed
    if n_qubits < needed:
        raise ValueError(message)
    return n_qubits
