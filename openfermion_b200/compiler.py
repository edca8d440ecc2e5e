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


def resolve_n_qubits(operator, n_qubits, message='Invalid number of qubits specified.'):
    needed = count_qubits(operator)
    if n_qubits is None:
        return needed
    if n_qubits < needed:
        raise ValueError(message)
    return n_qubits
