"""Array-form fermion-to-qubit transforms.

The reference's `jordan_wigner` and `bravyi_kitaev` (transforms/opconversions/jordan_wigner.py:
23-331, bravyi_kitaev.py:19-716) multiply symbolic QubitOperators term by term in Python
(4.2 s for the 88 495-term 28-qubit molecular Hamiltonian, quartic in the number of orbitals).
Here each ladder operator is a pair of Pauli strings held as (x, z) bit masks
(hamiltonians.jw_tables / bk_tables), products of k ladder operators are expanded for all
terms at once with numpy mask arithmetic, and equal strings are merged with one sort.  The
result is a `hamiltonians.MaskOperator`: the term compiler consumes its arrays directly, and
`.terms` gives the reference's dict when a caller wants one.

Same operator as the reference's output (coefficients equal to rounding, small sums dropped
with the reference's 1e-8 rule); the term ORDER is by (x, z) masks instead of insertion order.
"""
import numpy

from openfermion_b200 import hamiltonians, ops


def _fermion_term_list(operator):
    """[(modes, kinds, coefficient)] of a FermionOperator-like object."""
    out = []
    for term, coefficient in operator.terms.items():
        modes = tuple(int(mode) for mode, _ in term)
        kinds = tuple(int(kind) for _, kind in term)
        out.append((modes, kinds, complex(coefficient)))
    return out


def _count_modes(operator):
    return ops.count_qubits(operator)


def _transform(operator, n_qubits, encoding, device=False):
    names = {c.__name__ for c in type(operator).__mro__}
    if hasattr(operator, 'one_body_tensor') and hasattr(operator, 'two_body_tensor'):
        needed = numpy.asarray(operator.one_body_tensor).shape[0]
        if n_qubits is not None and n_qubits < needed:
            raise ValueError('Invalid number of qubits specified.')
        return hamiltonians.interaction_operator_jw(operator.constant, operator.one_body_tensor,
                                                    operator.two_body_tensor, encoding=encoding,
                                                    n_qubits=n_qubits, device=device)
    if 'DiagonalCoulombHamiltonian' in names or (
            all(hasattr(operator, a) for a in ('one_body', 'two_body', 'constant'))
            and not hasattr(operator, 'terms')):
        from openfermion_b200.linalg.sparse_tools import _diagonal_coulomb_fermion_terms
        operator = _diagonal_coulomb_fermion_terms(operator)
    if not ops.is_fermion_operator(operator):
        raise TypeError('Operator must be a FermionOperator, DiagonalCoulombHamiltonian, or '
                        'InteractionOperator.')
    needed = _count_modes(operator)
    if n_qubits is None:
        n_qubits = needed
    if n_qubits < needed:
        raise ValueError('Invalid number of qubits specified.')
    return hamiltonians._jw_of_term_list(_fermion_term_list(operator), n_qubits, encoding, device)


def jordan_wigner(operator, device=False):
    """Jordan-Wigner image of a FermionOperator / InteractionOperator-like tensor object /
    DiagonalCoulombHamiltonian-like object (jordan_wigner.py:23-65) as a MaskOperator.
    device=True expands the ladder products and merges equal strings on the GPU
    (csrc/termgen.cu): same arrays, bit for bit."""
    return _transform(operator, None, 'jw', device)


def bravyi_kitaev(operator, n_qubits=None, device=False):
    """Bravyi-Kitaev image (bravyi_kitaev.py:19-52, Fenwick-tree update / parity / occupation
    sets); n_qubits pads the tree like the reference's argument does; device as above."""
    return _transform(operator, n_qubits, 'bk', device)
