"""Operator files (openfermion_b200/operator_files.py) against files written by the reference's
save_operator (tests/golden/operator_file_*.data, oracle/make_golden.py operator_file_cases)
and round trips of both formats."""
import marshal
import os

import pytest

from helpers import GOLDEN
from openfermion_b200 import hamiltonians, operator_files, ops


@pytest.fixture(autouse=True)
def _stand_in_classes(monkeypatch):
    # exercise the repo's own term containers even where upstream OpenFermion is importable
    monkeypatch.setattr(operator_files, '_operator_classes',
                        lambda: {'FermionOperator': ops.FermionOperator, 'QubitOperator': ops.QubitOperator})


def test_reads_plain_text_files_written_by_the_reference():
    fop = operator_files.load_operator('operator_file_fermion', GOLDEN, plain_text=True)
    want = ops.FermionOperator()
    for modes, kinds, value in hamiltonians.hubbard_fermion_terms(2, 2, 1.0, 4.0, 0.3, 0.2):
        want += ops.FermionOperator(tuple(zip(modes, kinds)), value)
    assert isinstance(fop, ops.FermionOperator) and fop == want
    qop = operator_files.load_operator('operator_file_qubit.data', GOLDEN, plain_text=True)
    expected = hamiltonians.hubbard_jw(2, 2, 1.0, 4.0, 0.3, 0.2)
    terms = dict(expected.terms)
    terms[((0, 'X'), (3, 'Y'))] = 0.5 - 0.25j
    assert set(qop.terms) == set(terms)
    assert all(abs(qop.terms[k] - terms[k]) < 1e-12 for k in terms)


def test_plain_text_output_is_byte_identical_to_the_reference(tmp_path):
    for name in ('operator_file_fermion', 'operator_file_qubit'):
        op = operator_files.load_operator(name, GOLDEN, plain_text=True)
        operator_files.save_operator(op, name, str(tmp_path), plain_text=True)
        assert open(os.path.join(str(tmp_path), name + '.data')).read() == \
            open(os.path.join(GOLDEN, name + '.data')).read()


def test_marshal_round_trip_and_foreign_dump(tmp_path):
    op = ops.QubitOperator('X0 Z2', 1.5) + ops.QubitOperator('Y1', -0.25j) + ops.QubitOperator((), 2.0)
    operator_files.save_operator(op, 'm', str(tmp_path))
    assert operator_files.load_operator('m', str(tmp_path)) == op
    # the reference's binary layout: marshal of (type name, {term: complex})
    with open(os.path.join(str(tmp_path), 'raw.data'), 'wb') as f:
        marshal.dump(('FermionOperator', {((2, 1), (0, 0)): 0.5 + 0j, (): 1 + 0j}), f)
    fop = operator_files.load_operator('raw', str(tmp_path))
    assert fop.terms == {((2, 1), (0, 0)): 0.5 + 0j, (): 1 + 0j}


def test_errors_follow_reference(tmp_path):
    op = ops.QubitOperator('X0')
    with pytest.raises(operator_files.OperatorUtilsError, match='File name is not provided'):
        operator_files.save_operator(op, None, str(tmp_path))
    operator_files.save_operator(op, 'x', str(tmp_path))
    with pytest.raises(operator_files.OperatorUtilsError, match='already exists'):
        operator_files.save_operator(op, 'x', str(tmp_path))
    operator_files.save_operator(op, 'x', str(tmp_path), allow_overwrite=True)
    with pytest.raises(TypeError, match='Operator of invalid type'):
        operator_files.save_operator('hello', 'y', str(tmp_path))
    with open(os.path.join(str(tmp_path), 'bad.data'), 'wb') as f:
        marshal.dump(('str', 'hello'), f)  # the reference's bad_type_operator.data
    with pytest.raises(TypeError, match='Operator of invalid type'):
        operator_files.load_operator('bad', str(tmp_path))
    with open(os.path.join(str(tmp_path), 'boson.data'), 'w') as f:
        f.write('BosonOperator:\n1.0 [0^ 1]')
    with pytest.raises(TypeError):
        operator_files.load_operator('boson', str(tmp_path), plain_text=True)
