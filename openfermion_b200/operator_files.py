"""On-disk operator files of the reference (second half).

`save_operator` / `load_operator` (utils/operator_utils.py:221-360) store a QubitOperator or
FermionOperator either as a `marshal` dump of `(type name, {term: complex})` or as plain text
`"<type name>:\\n<coeff> [<factors>] +\\n..."` (the reference's `str(operator)`).  The functions
below read and write both formats with the term containers of `openfermion_b200.ops` (or the
upstream classes when OpenFermion is importable), so saved Hamiltonians feed the GPU path
without the symbolic layer.  Boson / quadrature operators are outside the Pauli-sum path and
raise TypeError.
"""
import marshal
import os
import re

from openfermion_b200 import ops

_TERM_PATTERN = re.compile(r'([^[\]]*)\[([^\]]*)\]', flags=re.DOTALL)


class OperatorUtilsError(Exception):
    pass


def get_file_path(file_name, data_directory):
    """'<data_directory>/<file_name>.data' (operator_utils.py:221-245); the reference's default
    directory is its package data folder, here the current directory."""
    if file_name:
        if file_name[-5:] != '.data':
            file_name = file_name + '.data'
    else:
        raise OperatorUtilsError('File name is not provided.')
    return (data_directory if data_directory is not None else os.getcwd()) + '/' + file_name


def _operator_classes():
    try:  # upstream classes when OpenFermion itself is installed
        from openfermion.ops import FermionOperator, QubitOperator
        return {'FermionOperator': FermionOperator, 'QubitOperator': QubitOperator}
    except ImportError:
        return {'FermionOperator': ops.FermionOperator, 'QubitOperator': ops.QubitOperator}


def _parse_coefficient(text):
    # symbolic_operator.py:168-186
    text = re.sub(r'\s+', '', text)
    if text and text[0] == '+':
        text = text[1:].strip()
    if text == '':
        return 1.0
    if text == '-':
        return -1.0
    try:
        if 'j' in text:
            return -complex(text[1:]) if text[0] == '-' else complex(text)
        return float(text)
    except ValueError:
        raise ValueError('Invalid coefficient {}.'.format(text))


def _parse_long_string(cls, long_string):
    """'1.5 [2^ 3] + 1.4 [3^ 0]' -> operator (symbolic_operator.py:158-195); every bracket is
    handed to the class's own single-term parser, equal terms are summed."""
    operator = cls()
    for coefficient_text, factors in _TERM_PATTERN.findall(long_string):
        coefficient = _parse_coefficient(coefficient_text)
        piece = cls(factors.strip() or (), coefficient) if factors.strip() else cls((), coefficient)
        for term, value in piece.terms.items():
            if term not in operator.terms:
                operator.terms[term] = value
            else:
                operator.terms[term] += value
    return operator


def load_operator(file_name=None, data_directory=None, plain_text=False):
    """Stored QubitOperator or FermionOperator (operator_utils.py:248-306)."""
    file_path = get_file_path(file_name, data_directory)
    classes = _operator_classes()
    if plain_text:
        with open(file_path, 'r') as f:
            operator_type, operator_terms = f.read().split(':\n')
        if operator_type in ('BosonOperator', 'QuadOperator') or operator_type not in classes:
            raise TypeError('Operator of invalid type.')
        return _parse_long_string(classes[operator_type], operator_terms)
    with open(file_path, 'rb') as f:
        data = marshal.load(f)
    operator_type, operator_terms = data[0], data[1]
    if operator_type not in classes:
        raise TypeError('Operator of invalid type.')
    cls = classes[operator_type]
    operator = cls()
    for term in operator_terms:
        operator += cls(term, operator_terms[term])
    return operator


def _term_string(operator, term):
    if ops.is_qubit_operator(operator):
        return ' '.join('{}{}'.format(action, index) for index, action in term)
    return ' '.join('{}{}'.format(index, '^' if action else '') for index, action in term)


def operator_string(operator):
    """str(operator) of the reference (symbolic_operator.py:334-351): terms sorted, small
    coefficients skipped, one '<coeff> [<factors>] +' line per term."""
    if not operator.terms:
        return '0'
    lines = []
    for term, coefficient in sorted(operator.terms.items()):
        if abs(coefficient) < ops.EQ_TOLERANCE:
            continue
        lines.append('{} [{}] +\n'.format(coefficient, _term_string(operator, term)))
    return ''.join(lines)[:-3]


def save_operator(operator, file_name=None, data_directory=None, allow_overwrite=False, plain_text=False):
    """Writes the reference's file formats (operator_utils.py:309-360)."""
    file_path = get_file_path(file_name, data_directory)
    if os.path.isfile(file_path) and not allow_overwrite:
        raise OperatorUtilsError('Not saved, file already exists.')
    if ops.is_fermion_operator(operator):
        operator_type = 'FermionOperator'
    elif ops.is_qubit_operator(operator):
        operator_type = 'QubitOperator'
    else:
        raise TypeError('Operator of invalid type.')
    for coefficient in operator.terms.values():
        if type(coefficient).__module__.startswith('sympy'):
            raise TypeError('Cannot save sympy expressions.')
    if plain_text:
        with open(file_path, 'w') as f:
            f.write(operator_type + ':\n' + operator_string(operator))
    else:
        terms = operator.terms
        with open(file_path, 'wb') as f:
            marshal.dump((operator_type, dict(zip(terms.keys(), map(complex, terms.values())))), f)
