"""Host logic of the symmetry-adapted sharding (openfermion_b200/shard_basis.py): GF(2)
algebra, the choice of shard rows, and -- against the oracle's restatement of the reference
matvec (linalg/linear_qubit_operator.py:107-148) -- that the relabelled term table is the
same operator on the relabelled state."""
import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import hamiltonians
from openfermion_b200.compiler import compile_qubit_operator
from openfermion_b200.distributed import ShardLayout
from openfermion_b200.ops import QubitOperator
from openfermion_b200.shard_basis import ShardBasis, _distinct_outside, gf2_nullspace, locality_order


def _gather_apply(n, x, z, c, v):
    """Gather form over (x, z, Pauli coefficient) rows: the formula the kernels implement."""
    j = numpy.arange(1 << n, dtype=numpy.uint64)
    out = numpy.zeros(1 << n, dtype=numpy.complex128)
    for t in range(len(c)):
        y = bin(int(x[t] & z[t])).count('1')
        sign = 1.0 - 2.0 * po.bit_parity(j & z[t]).astype(float)
        out += c[t] * (-1j) ** y * sign * v[(j ^ x[t]).astype(numpy.int64)]
    return out


def test_nullspace_is_orthogonal_and_complete():
    rng = numpy.random.default_rng(0)
    for n, count in ((6, 3), (10, 7), (17, 20), (40, 25)):
        masks = [int(rng.integers(0, 1 << min(n, 62))) for _ in range(count)]
        basis = gf2_nullspace(masks, n)
        for m in basis:
            assert all(bin(m & x).count('1') % 2 == 0 for x in masks)
        # rank-nullity: independent vectors, dimension n - rank
        echelon = {}
        rank = 0
        for x in masks:
            while x and (x.bit_length() - 1) in echelon:
                x ^= echelon[x.bit_length() - 1]
            if x:
                echelon[x.bit_length() - 1] = x
                rank += 1
        indep = {}
        for m in basis:
            while m and (m.bit_length() - 1) in indep:
                m ^= indep[m.bit_length() - 1]
            assert m
            indep[m.bit_length() - 1] = m
        assert len(basis) == n - rank


def test_index_maps_are_inverse_permutations():
    rng = numpy.random.default_rng(1)
    n = 9
    rows = [1 << b for b in range(n)]
    for _ in range(40):  # random invertible matrix from row operations
        a, b = rng.integers(0, n, size=2)
        if a != b:
            rows[a] ^= rows[b]
    basis = ShardBasis(n, rows)
    idx = numpy.arange(1 << n, dtype=numpy.uint64)
    fwd = basis.to_basis_index(idx)
    assert sorted(fwd.tolist()) == idx.tolist()
    assert numpy.array_equal(basis.to_natural_index(fwd), idx)
    v = rng.normal(size=1 << n)
    assert numpy.array_equal(basis.to_natural(basis.from_natural(v)), v)
    assert basis.from_natural(v)[int(fwd[37])] == v[37]


def test_singular_basis_is_rejected():
    with pytest.raises(ValueError):
        ShardBasis(3, [1, 1, 4])
    with pytest.raises(ValueError):
        ShardBasis(3, [1, 2])


@pytest.mark.parametrize('local_order', ['natural', 'locality'])
@pytest.mark.parametrize('kind', ['hubbard', 'hubbard_bk', 'molecular', 'molecular_bk', 'complex', 'ys'])
@pytest.mark.parametrize('shard_bits', [0, 1, 2, 3])
def test_relabelled_table_is_the_same_operator(kind, shard_bits, local_order):
    if kind == 'hubbard':
        op, n = hamiltonians.hubbard_jw(2, 2, chemical_potential=0.3, magnetic_field=0.2), 8
    elif kind == 'hubbard_bk':
        op, n = hamiltonians.hubbard_jw(2, 2, encoding='bk'), 8
    elif kind == 'molecular':
        op, n = hamiltonians.molecular_jw(4, seed=7), 8
    elif kind == 'molecular_bk':
        op, n = hamiltonians.molecular_jw(4, seed=7, encoding='bk'), 8
    elif kind == 'complex':
        op, n = hamiltonians.random_pauli_sum(8, 60, 3, True), 8
    else:  # strings with 1..4 Y factors: every residue of the i^y phase
        op = QubitOperator('Y0', 1.5) + QubitOperator('Y0 Y1', -0.5) + QubitOperator('Y0 Y3 Y5', 2.0) \
            + QubitOperator('Y1 Y2 Y4 Y6', 0.75j) + QubitOperator('X0 Y2 Z3', 1.0) + QubitOperator('Z1 Z6', 3.0)
        n = 7
    x, z, c = compile_qubit_operator(op, n)
    basis = ShardBasis.for_sharding(n, shard_bits, x, local_order=local_order)
    xp, zp, cp = basis.transform_terms(x, z, c)
    rng = numpy.random.default_rng(5)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    want = po.matvec(op.terms, n, v)
    got = basis.to_natural(_gather_apply(n, xp, zp, cp, basis.from_natural(v)))
    assert numpy.max(numpy.abs(got - want)) <= 1e-13 * numpy.max(numpy.abs(want))
    # never more peer shards than sharding by the top qubits
    world = 1 << shard_bits
    assert len(ShardLayout(n, world, xp).remote_patterns) <= len(ShardLayout(n, world, x).remote_patterns)
    # the non-shard bits are single qubits; in their natural order unless re-ordered for locality
    local_rows = basis.rows[:n - shard_bits]
    assert all(bin(r).count('1') == 1 for r in local_rows)
    if local_order == 'natural':
        assert local_rows == sorted(local_rows)
    else:
        assert basis.rows[n - shard_bits:] == ShardBasis.for_sharding(n, shard_bits, x).rows[n - shard_bits:]


def test_locality_order_reduces_distinct_gather_blocks():
    """The qubit order inside a shard is chosen to minimise the number of distinct x-mask parts
    above the L1 tile (2^10 amplitudes) and above the L2 span (2^19): the model that matches the
    measured L1 miss share and DRAM streams of the Hubbard runs (profiles/README.md)."""
    op = hamiltonians.hubbard_jw(4, 4)
    x, z, c = compile_qubit_operator(op, 32)
    for shard_bits, natural_costs in ((0, (51, 30)), (3, (43, 19))):
        local = 32 - shard_bits
        costs = []
        for local_order in ('natural', 'locality'):
            basis = ShardBasis.for_sharding(32, shard_bits, x, local_order=local_order)
            xp, _, _ = basis.transform_terms(x, z, c)
            g = numpy.unique(xp & numpy.uint64((1 << local) - 1))
            costs.append((_distinct_outside(g, (1 << 10) - 1), _distinct_outside(g, (1 << 19) - 1)))
        assert costs[0] == natural_costs
        assert costs[1][0] <= 0.8 * costs[0][0] and costs[1][1] <= 0.65 * costs[0][1]
    # deterministic, a permutation, and the identity when nothing can be gained
    order = locality_order(numpy.unique(x), 32)
    assert order == locality_order(numpy.unique(x), 32) and sorted(order) == list(range(32))
    assert locality_order(numpy.array([0, 1, 2, 3], dtype=numpy.uint64), 12) == list(range(12))
    dense = hamiltonians.molecular_jw(10, seed=1234)  # 2 536 groups: nothing to gain, natural order kept
    xd, _ = dense.index_masks(20)
    assert ShardBasis.for_sharding(20, 0, xd, local_order='locality').identity


def test_conserved_parities_remove_the_exchange():
    """SURVEY.md section 8e counts 1 / 2 / 4 peer shards per H|psi> for the 4x4 Hubbard model on
    2 / 4 / 8 ranks when sharding by the top qubits; the spin-parity shard bits need 0 / 0 / 1."""
    op = hamiltonians.hubbard_jw(4, 4)
    x, z, c = compile_qubit_operator(op, 32)
    for world, natural, adapted in ((2, 1, 0), (4, 2, 0), (8, 4, 1), (16, 6, 2)):
        k = world.bit_length() - 1
        basis = ShardBasis.for_sharding(32, k, x)
        xp, zp, cp = basis.transform_terms(x, z, c)
        assert len(ShardLayout(32, world, x).remote_patterns) == natural
        layout = ShardLayout(32, world, xp)
        assert len(layout.remote_patterns) == adapted
        assert len(layout.group_x) == 65 and len(cp) == len(c)
        # up-spin qubits are the even ones: index bits 31, 29, ...
        up = sum(1 << (31 - q) for q in range(0, 32, 2))
        assert basis.rows[31] in (up, up >> 1)
    mol = hamiltonians.molecular_jw(10, seed=1234)
    x, z, c = compile_qubit_operator(mol, 20)
    xp, _, _ = ShardBasis.for_sharding(20, 3, x).transform_terms(x, z, c)
    assert len(ShardLayout(20, 8, x).remote_patterns) == 7
    assert len(ShardLayout(20, 8, xp).remote_patterns) == 1


def test_natural_basis_is_a_no_op():
    op = hamiltonians.random_pauli_sum(6, 20, 1, True)
    x, z, c = compile_qubit_operator(op, 6)
    basis = ShardBasis.natural(6)
    assert basis.identity
    xp, zp, cp = basis.transform_terms(x, z, c)
    assert numpy.array_equal(xp, x) and numpy.array_equal(zp, z) and numpy.array_equal(cp, c)
    assert numpy.array_equal(basis.shard_natural_indices(2, 4), numpy.arange(32, 48, dtype=numpy.uint64))
    assert ShardBasis.for_sharding(6, 0, x).identity


@pytest.mark.parametrize('seed', range(6))
def test_general_invertible_relabelling(seed):
    """Any invertible binary matrix is a valid basis (the `basis=ShardBasis(...)` option): dense
    random matrices, strings with every Y-count residue, complex coefficients."""
    rng = numpy.random.default_rng(100 + seed)
    n = 7
    rows = [1 << b for b in range(n)]
    for _ in range(60):
        a, b = (int(v) for v in rng.integers(0, n, size=2))
        if a != b:
            rows[a] ^= rows[b]
        if rng.random() < 0.3:
            a, b = (int(v) for v in rng.integers(0, n, size=2))
            rows[a], rows[b] = rows[b], rows[a]
    basis = ShardBasis(n, rows)
    op = hamiltonians.random_pauli_sum(n, 50, seed, True)
    x, z, c = compile_qubit_operator(op, n)
    xp, zp, cp = basis.transform_terms(x, z, c)
    # a relabelling maps distinct strings to distinct strings and keeps |coefficients|
    assert len({(int(a), int(b)) for a, b in zip(xp, zp)}) == len({(int(a), int(b)) for a, b in zip(x, z)})
    assert numpy.array_equal(numpy.abs(cp), numpy.abs(c))
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    want = po.matvec(op.terms, n, v)
    got = basis.to_natural(_gather_apply(n, xp, zp, cp, basis.from_natural(v)))
    assert numpy.max(numpy.abs(got - want)) <= 1e-13 * numpy.max(numpy.abs(want))
    # composing with the inverse relabelling gives the original table back
    inverse = ShardBasis(n, basis.inv_rows)
    xb, zb, cb = inverse.transform_terms(xp, zp, cp)
    assert numpy.array_equal(xb, x) and numpy.array_equal(zb, z) and numpy.array_equal(cb, c)
