"""Minimal term containers so the hot path can be fed without upstream OpenFermion.

The hot path only reads `operator.terms` (ops/operators/symbolic_operator.py:58-65 in the
reference): a dict {((index, action), ...): coefficient}.  Upstream `QubitOperator` /
`FermionOperator` instances are accepted everywhere; these two small classes exist for
machines where upstream is not installed (the GPU box) and implement only what the
linalg functions, tests and benches need: construction from the reference's string /
tuple syntax, +, -, scalar *, Pauli products, compress() and get_operator_groups().
The symbolic algebra of the reference (sympy coefficients, ...) is out of scope
(SURVEY.md section 2, rows 5 and 7); fermionic normal ordering is restated below because
get_number_preserving_sparse_operator sums its terms in normal-ordered term order.
"""
import re

EQ_TOLERANCE = 1e-8  # reference config.py:17

_PAULI_X = {'I': 0, 'X': 1, 'Y': 1, 'Z': 0}
_PAULI_Z = {'I': 0, 'X': 0, 'Y': 1, 'Z': 1}
_FROM_XZ = {(0, 0): 'I', (1, 0): 'X', (1, 1): 'Y', (0, 1): 'Z'}
_I_POW = (1, 1j, -1, -1j)


class _TermDict:
    """Shared behaviour: a linear combination of keyed terms."""

    def __init__(self):
        self.terms = {}

    def _add_term(self, key, coefficient):
        # same rule as the reference's += (symbolic_operator.py:433-456): sums whose
        # magnitude falls below EQ_TOLERANCE are removed from the dict
        # -- whether or not the key was there before: a tiny new term never enters the dict
        total = self.terms.get(key, 0) + coefficient
        if abs(total) < EQ_TOLERANCE:
            self.terms.pop(key, None)
        else:
            self.terms[key] = total

    def __iadd__(self, other):
        if not isinstance(other, type(self)):
            raise TypeError('Cannot add {} to {}'.format(type(other).__name__, type(self).__name__))
        for key, coefficient in other.terms.items():
            self._add_term(key, coefficient)
        return self

    def __add__(self, other):
        out = self._copy()
        out += other
        return out

    def __neg__(self):
        return self * -1.0

    def __isub__(self, other):
        self += other * -1.0
        return self

    def __sub__(self, other):
        out = self._copy()
        out -= other
        return out

    def _copy(self):
        out = type(self)()
        out.terms = dict(self.terms)
        return out

    def __rmul__(self, scalar):
        return self * scalar

    def __truediv__(self, scalar):
        return self * (1.0 / scalar)

    def compress(self, abs_tol=EQ_TOLERANCE):
        """Drops small terms and casts numerically real coefficients to float
        (symbolic_operator.py:694-725)."""
        new_terms = {}
        for key, c in self.terms.items():
            if isinstance(c, complex) or hasattr(c, 'imag'):
                if abs(complex(c).imag) <= abs_tol:
                    c = complex(c).real
                elif abs(complex(c).real) <= abs_tol:
                    c = 1j * complex(c).imag
            if abs(c) > abs_tol:
                new_terms[key] = c
        self.terms = new_terms

    def get_operator_groups(self, num_groups):
        """Contiguous slices of the term dict (symbolic_operator.py:779-797)."""
        items = list(self.terms.items())
        n = len(items)
        if n == 0:
            return
        num_groups = max(1, min(num_groups, n))
        for g in range(num_groups):
            lo, hi = n * g // num_groups, n * (g + 1) // num_groups
            if hi > lo:
                piece = type(self)()
                piece.terms = dict(items[lo:hi])
                yield piece

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        keys = set(self.terms) | set(other.terms)
        return all(abs(self.terms.get(k, 0) - other.terms.get(k, 0)) <= EQ_TOLERANCE for k in keys)

    def __repr__(self):
        return '{}({} terms)'.format(type(self).__name__, len(self.terms))


class QubitOperator(_TermDict):
    """Sum of Pauli strings; term keys are ((qubit, 'X'|'Y'|'Z'), ...) sorted by qubit."""

    actions = ('X', 'Y', 'Z')

    def __init__(self, term=None, coefficient=1.0):
        super().__init__()
        if term is None:
            return
        factors = self._parse(term)
        key, phase = self._simplify(factors)
        self.terms[key] = coefficient * phase if phase != 1 else coefficient

    @staticmethod
    def _parse(term):
        if isinstance(term, str):
            factors = []
            for token in term.split():
                m = re.fullmatch(r'([XYZ])(\d+)', token)
                if not m:
                    raise ValueError('Invalid Pauli factor {!r}'.format(token))
                factors.append((int(m.group(2)), m.group(1)))
            return factors
        factors = list(term)
        if factors and not isinstance(factors[0], (tuple, list)):
            factors = [tuple(factors)]
        for q, a in factors:
            if a not in ('X', 'Y', 'Z') or int(q) < 0:
                raise ValueError('Invalid Pauli factor {!r}'.format((q, a)))
        return [(int(q), a) for q, a in factors]

    @staticmethod
    def _simplify(factors):
        """Multiplies factors on equal qubits left to right (qubit_operator.py:116-145)."""
        x, z, power = {}, {}, 0
        for q, a in factors:
            ax, az = _PAULI_X[a], _PAULI_Z[a]
            bx, bz = x.get(q, 0), z.get(q, 0)
            # current * new with P(x,z) = i^{xz} X^x Z^z
            power += bx * bz + ax * az + 2 * (bz * ax)
            x[q], z[q] = bx ^ ax, bz ^ az
            power -= x[q] * z[q]
        key = tuple((q, _FROM_XZ[(x[q], z[q])]) for q in sorted(x) if (x[q], z[q]) != (0, 0))
        return key, _I_POW[power % 4]

    def __mul__(self, other):
        if isinstance(other, QubitOperator):
            out = QubitOperator()
            for ka, ca in self.terms.items():
                for kb, cb in other.terms.items():
                    key, phase = self._simplify(list(ka) + list(kb))
                    out._add_term(key, ca * cb * phase)
            return out
        out = self._copy()
        for k in out.terms:
            out.terms[k] = out.terms[k] * other
        return out

    def __imul__(self, other):
        self.terms = (self * other).terms
        return self


class FermionOperator(_TermDict):
    """Sum of ladder-operator products; keys are ((mode, 1|0), ...), 1 = raising."""

    actions = (1, 0)

    def __init__(self, term=None, coefficient=1.0):
        super().__init__()
        if term is None:
            return
        if isinstance(term, str):
            factors = []
            for token in term.split():
                m = re.fullmatch(r'(\d+)(\^?)', token)
                if not m:
                    raise ValueError('Invalid ladder factor {!r}'.format(token))
                factors.append((int(m.group(1)), 1 if m.group(2) else 0))
            key = tuple(factors)
        else:
            factors = list(term)
            if factors and not isinstance(factors[0], (tuple, list)):
                factors = [tuple(factors)]
            key = tuple((int(q), int(a)) for q, a in factors)
        self.terms[key] = coefficient

    def __mul__(self, other):
        if isinstance(other, FermionOperator):
            out = FermionOperator()
            for ka, ca in self.terms.items():
                for kb, cb in other.terms.items():
                    out._add_term(ka + kb, ca * cb)
            return out
        out = self._copy()
        for k in out.terms:
            out.terms[k] = out.terms[k] * other
        return out


def _normal_ordered_ladder_term(term, coefficient):
    """Normal-ordered form of one fermionic ladder product: raising operators first, modes
    descending (transforms/opconversions/term_reordering.py:153-233 with parity = -1)."""
    term = list(term)
    ordered = FermionOperator()
    for i in range(1, len(term)):
        for j in range(i, 0, -1):
            right, left = term[j], term[j - 1]
            if right[1] and not left[1]:
                # a a^ = delta - a^ a
                term[j - 1], term[j] = right, left
                coefficient = coefficient * -1
                if right[0] == left[0]:
                    ordered += _normal_ordered_ladder_term(tuple(term[:j - 1] + term[j + 1:]),
                                                           -1 * coefficient)
            elif right[1] == left[1]:
                if right[0] == left[0]:
                    return ordered  # a a = a^ a^ = 0
                if right[0] > left[0]:
                    term[j - 1], term[j] = right, left
                    coefficient = coefficient * -1
    ordered += FermionOperator(tuple(term), coefficient)
    return ordered


def normal_ordered(operator):
    """normal_ordered for FermionOperators (term_reordering.py:62-150, fermionic branch):
    terms are processed in dict order and accumulated with +=."""
    if not is_fermion_operator(operator):
        raise TypeError('Can only normal order FermionOperator here.')
    ordered = FermionOperator()
    for term, coefficient in operator.terms.items():
        ordered += _normal_ordered_ladder_term(term, coefficient)
    return ordered


def is_normal_ordered(operator):
    """FermionOperator.is_normal_ordered (ops/operators/fermion_operator.py): in every term the
    raising operators stand left of the lowering ones and indices descend within each kind."""
    for term in operator.terms:
        for left, right in zip(term, term[1:]):
            if right[1] and not left[1]:
                return False
            if right[1] == left[1] and right[0] >= left[0]:
                return False
    return True


def _class_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def is_qubit_operator(obj):
    return ('QubitOperator' in _class_names(obj) or hasattr(obj, 'index_masks')) and hasattr(obj, 'terms')


def is_fermion_operator(obj):
    return 'FermionOperator' in _class_names(obj) and hasattr(obj, 'terms')


def count_qubits(operator):
    """1 + highest index appearing in the terms (utils/operator_utils.py:143-202)."""
    if hasattr(operator, 'index_masks'):  # hamiltonians.MaskOperator: array form
        return operator.count_qubits()
    if not hasattr(operator, 'terms'):
        raise TypeError('Operator of invalid type.')
    n = 0
    for term in operator.terms:
        for factor in term:
            if factor[0] + 1 > n:
                n = factor[0] + 1
    return n
