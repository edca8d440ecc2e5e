"""Support hints (openfermion_b200/support.py, ofb_plan_set_support): the conditions derived from
the term table really imply S_g(j) = 0, and the bit tricks the kernel uses to turn a condition
into the 16-bit "needed amplitudes" mask of a thread (csrc/apply.cu, SegCond) are restated here
and checked against the definition."""
import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import hamiltonians, support
from openfermion_b200.compiler import compile_qubit_operator
from openfermion_b200.shard_basis import ShardBasis

CLASS_SHIFT, CLASS_BITS = 6, 4  # OFB_APPLY_BLOCK_BITS / OFB_APPLY_CLASS_BITS of the shipped build


def _coefficient_function(x, z, k, xg, n):
    j = numpy.arange(1 << n, dtype=numpy.uint64)
    s = numpy.zeros(1 << n, dtype=numpy.complex128)
    for t in numpy.nonzero(x == xg)[0]:
        s += k[t] * (1.0 - 2.0 * po.bit_parity(j & z[t]))
    return s


CASES = [
    ('hubbard_2x2', lambda: hamiltonians.hubbard_jw(2, 2), 8, 0.56),
    ('hubbard_2x3_mu_h', lambda: hamiltonians.hubbard_jw(2, 3, 1.0, 4.0, 0.3, 0.2), 12, 0.53),
    ('hubbard_2x2_bk', lambda: hamiltonians.hubbard_jw(2, 2, encoding='bk'), 8, 0.56),
    ('molecular_4', lambda: hamiltonians.molecular_jw(4, seed=3), 8, 0.34),
    ('molecular_5_bk', lambda: hamiltonians.molecular_jw(5, seed=3, encoding='bk'), 10, 0.32),
    ('random_paulis', lambda: hamiltonians.random_pauli_sum(8, 60, 1, True), 8, 1.0),
]


@pytest.mark.parametrize('name,builder,n,fraction', CASES, ids=[c[0] for c in CASES])
def test_conditions_imply_zero_coefficients(name, builder, n, fraction):
    op = builder()
    x, z, c = compile_qubit_operator(op, n)
    group_x, mask1, mask2, flags = support.group_conditions(x, z, c)
    assert numpy.array_equal(group_x, numpy.unique(x))
    k = support._phased(x, z, c)
    j = numpy.arange(1 << n, dtype=numpy.uint64)
    for g, xg in enumerate(group_x):
        s = _coefficient_function(x, z, k, xg, n)
        need = numpy.ones(1 << n, dtype=bool)
        if flags[g] & 1:
            need &= po.bit_parity(j & mask1[g]) == ((flags[g] >> 1) & 1)
        if flags[g] & 4:
            need &= po.bit_parity(j & mask2[g]) == ((flags[g] >> 3) & 1)
            assert mask2[g] != mask1[g] and mask2[g] != 0
        scale = numpy.max(numpy.abs(k[x == xg]))
        assert numpy.all(numpy.abs(s[~need]) <= 1e-13 * scale)
        if xg == 0:
            assert flags[g] == 0
    assert abs(support.needed_fraction(flags) - fraction) < 0.01


@pytest.mark.parametrize('name,builder,n,fraction', CASES, ids=[c[0] for c in CASES])
def test_reduced_table_equals_the_operator_on_the_support(name, builder, n, fraction):
    """One term per symmetry orbit (support.reduce_terms) + zeroing outside the support -- what
    the hinted kernel computes -- is the reference matvec."""
    op = builder()
    x, z, c = compile_qubit_operator(op, n)
    group_x, mask1, mask2, flags = support.group_conditions(x, z, c)
    xr, zr, cr = support.reduce_terms(x, z, c, group_x, mask1, mask2, flags)
    assert numpy.array_equal(numpy.unique(xr), group_x)
    assert len(cr) <= len(c) and (len(cr) < len(c)) == bool(numpy.any(flags))
    kr = support._phased(xr, zr, cr)
    rng = numpy.random.default_rng(2)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    j = numpy.arange(1 << n, dtype=numpy.uint64)
    out = numpy.zeros(1 << n, dtype=numpy.complex128)
    for g, xg in enumerate(group_x):
        need = numpy.ones(1 << n, dtype=bool)
        if flags[g] & 1:
            need &= po.bit_parity(j & mask1[g]) == ((flags[g] >> 1) & 1)
        if flags[g] & 4:
            need &= po.bit_parity(j & mask2[g]) == ((flags[g] >> 3) & 1)
        s = _coefficient_function(xr, zr, kr, xg, n)
        out += numpy.where(need, s, 0.0) * v[(j ^ xg).astype(numpy.int64)]
    want = po.matvec(op.terms, n, v)
    assert numpy.max(numpy.abs(out - want)) <= 1e-13 * numpy.max(numpy.abs(want))


def test_conditions_survive_the_shard_relabelling():
    """The finder runs on whatever table the plan is built from: after the GF(2) relabelling of
    shard_basis.py the transformed strings have the transformed symmetry."""
    op = hamiltonians.hubbard_jw(2, 3)
    x, z, c = compile_qubit_operator(op, 12)
    basis = ShardBasis.for_sharding(12, 3, x, local_order='locality')
    xp, zp, cp = basis.transform_terms(x, z, c)
    _, _, _, flags = support.group_conditions(x, z, c)
    group_x, mask1, mask2, flags_p = support.group_conditions(xp, zp, cp)
    assert support.needed_fraction(flags_p) == support.needed_fraction(flags)
    k = support._phased(xp, zp, cp)
    j = numpy.arange(1 << 12, dtype=numpy.uint64)
    for g, xg in enumerate(group_x):
        if flags_p[g] & 1:
            s = _coefficient_function(xp, zp, k, xg, 12)
            off = po.bit_parity(j & mask1[g]) != ((flags_p[g] >> 1) & 1)
            assert numpy.all(s[off] == 0)


def test_full_size_tables():
    # the bench workloads: half of the Hubbard gathers and 70 % of the molecular ones are skippable
    hub = hamiltonians.hubbard_jw(4, 4)
    _, _, _, flags = support.group_conditions(*compile_qubit_operator(hub, 32))
    assert len(flags) == 65 and numpy.count_nonzero(flags & 1) == 64
    assert abs(support.needed_fraction(flags) - 33.0 / 65.0) < 1e-12
    mol = hamiltonians.molecular_jw(10, seed=1234)
    table = compile_qubit_operator(mol, 20)
    group_x, mask1, mask2, flags = support.group_conditions(*table)
    assert support.needed_fraction(flags) < 0.32
    xr, zr, cr = support.reduce_terms(*table, group_x, mask1, mask2, flags)
    assert len(cr) < 0.34 * len(table[2])  # 22 351 strings -> a third


def _seg_cond(mask1, mask2, flag):
    """SegCond packing of ofb_plan_set_support (csrc/plan.cu)."""
    def pattern(m):
        w = (m >> CLASS_SHIFT) & ((1 << CLASS_BITS) - 1)
        return sum(1 << r for r in range(1 << CLASS_BITS) if bin(r & w).count('1') & 1)
    full = (1 << (1 << CLASS_BITS)) - 1
    m1 = m2 = 0
    pat, v = full | (full << 16), 3
    if flag & 1:
        m1, pat, v = mask1, (pat & 0xffff0000) | pattern(mask1), (v & ~1) | ((flag >> 1) & 1)
    if flag & 4:
        m2, pat, v = mask2, (pat & 0x0000ffff) | (pattern(mask2) << 16), (v & ~2) | (((flag >> 3) & 1) << 1)
    return m1, m2, pat, v


def _kernel_need(j_lo, cond):
    """The need mask of k_apply_tile<..., COND = true> for the thread whose r = 0 amplitude is j_lo."""
    m1, m2, pat, v = cond
    q1 = (bin(j_lo & m1).count('1') + v) & 1
    q2 = (bin(j_lo & m2).count('1') + (v >> 1)) & 1
    return (pat ^ ((q1 - 1) & 0xffffffff)) & ((pat >> 16) ^ ((q2 - 1) & 0xffffffff)) & 0xffff


def test_kernel_need_mask_equals_the_definition():
    rng = numpy.random.default_rng(4)
    n = 20
    for trial in range(300):
        mask1 = int(rng.integers(1, 1 << n))
        mask2 = int(rng.integers(1, 1 << n))
        flag = int(rng.integers(0, 16)) & (0b1111 if trial % 3 else 0b0011)
        cond = _seg_cond(mask1, mask2, flag)
        tile = int(rng.integers(0, 1 << (n - 10))) << 10
        index_base = int(rng.integers(0, 4)) << n  # high bits of a shard's global index
        mask1_g, mask2_g = mask1 | (index_base & (1 << n)), mask2
        cond = _seg_cond(mask1_g, mask2_g, flag)
        for tid in rng.integers(0, 64, size=8):
            j_lo = index_base | tile | int(tid)
            need = _kernel_need(j_lo, cond)
            for r in range(16):
                j = j_lo ^ (r << CLASS_SHIFT)
                want = True
                if flag & 1:
                    want &= (bin(j & mask1_g).count('1') & 1) == ((flag >> 1) & 1)
                if flag & 4:
                    want &= (bin(j & mask2_g).count('1') & 1) == ((flag >> 3) & 1)
                assert bool((need >> r) & 1) == want


# ---- inside symmetry sectors ---------------------------------------------------------------
def _sector_emulation(n, x, z, c, tables, v):
    """What k_sector_apply computes: out[p] = sum_g [j ^ x_g member] S_g(j) in[pos(j ^ x_g)]."""
    k = support._phased(x, z, c)
    basis = tables.basis()
    out = numpy.zeros(tables.dim, dtype=numpy.complex128)
    for xg in numpy.unique(x):
        partner = basis ^ xg
        ok = tables.member(partner)
        s = numpy.zeros(tables.dim, dtype=numpy.complex128)
        for t in numpy.nonzero(x == xg)[0]:
            s += k[t] * (1.0 - 2.0 * po.bit_parity(basis & z[t]))
        pos = numpy.where(ok, tables.position(numpy.where(ok, partner, basis)), 0)
        out += numpy.where(ok, s * v[pos], 0.0)
    return out


@pytest.mark.parametrize('kind', ['sz_number', 'number', 'sz'])
@pytest.mark.parametrize('name,builder,n', [
    ('hubbard_2x3', lambda: hamiltonians.hubbard_jw(2, 3, 1.0, 4.0, 0.3, 0.2), 12),
    ('molecular_5', lambda: hamiltonians.molecular_jw(5, seed=4), 10),
])
def test_reduced_table_inside_a_sector(name, builder, n, kind):
    """Conditions implied by sector membership (support.conditions_implied_by_sector) let the
    unchanged sector kernel run on the orbit-reduced table.  With S_z and N fixed every
    condition of a spin-conserving Hamiltonian is implied; with N alone only the total-parity
    ones are, and the rest must stay unreduced."""
    from openfermion_b200 import sectors
    op = builder()
    x, z, c = compile_qubit_operator(op, n)
    if kind == 'sz_number':
        tables = sectors.sz_number_sector_tables(n, 0.0, 4)
    elif kind == 'number':
        tables = sectors.number_sector_tables(n, 4)
    else:
        tables = sectors.sz_sector_tables(n, 1.0)
    group_x, mask1, mask2, flags = support.group_conditions(x, z, c)
    implied = support.conditions_implied_by_sector(group_x, mask1, mask2, flags, tables.plus, tables.minus)
    assert numpy.all((implied & ~flags) == 0)
    xr, zr, cr = support.reduce_terms(x, z, c, group_x, mask1, mask2, implied)
    rng = numpy.random.default_rng(8)
    v = rng.normal(size=tables.dim) + 1j * rng.normal(size=tables.dim)
    want = _sector_emulation(n, x, z, c, tables, v)
    got = _sector_emulation(n, xr, zr, cr, tables, v)
    assert numpy.max(numpy.abs(got - want)) <= 1e-13 * numpy.max(numpy.abs(want))
    if kind == 'sz_number':
        assert numpy.array_equal(implied, flags) and len(cr) < 0.7 * len(c)
    if kind == 'number':  # mixed-spin double excitations keep their spin conditions unreduced
        assert len(c) >= len(cr)
        # using conditions that are NOT implied would be wrong: the check has teeth
        xb, zb, cb = support.reduce_terms(x, z, c, group_x, mask1, mask2, flags)
        bad = _sector_emulation(n, xb, zb, cb, tables, v)
        if len(cb) < len(cr):
            assert numpy.max(numpy.abs(bad - want)) > 1e-6 * numpy.max(numpy.abs(want))


def test_synthetic_symmetric_groups_with_complex_coefficients():
    """Groups built to be invariant under chosen (m, s) pairs -- complex coefficients, both signs,
    Y factors in the strings -- next to groups without any symmetry."""
    rng = numpy.random.default_rng(12)
    n = 9
    xs, zs, cs = [], [], []
    planted = {}
    for trial in range(12):
        xg = int(rng.integers(1, 1 << n))
        if xg in planted:
            continue
        gens = []
        for _ in range(int(rng.integers(0, 3))):
            m = int(rng.integers(1, 1 << n))
            if all(m != g[0] for g in gens) and (len(gens) < 1 or m != gens[0][0]):
                gens.append((m, float(rng.choice([-1.0, 1.0]))))
        elements = [(0, 1.0)]
        for m, s in gens:
            elements += [(e ^ m, se * s) for e, se in elements]
        if len({e for e, _ in elements}) != len(elements):
            continue  # dependent generators
        members = {}
        consistent = True
        for _ in range(int(rng.integers(1, 4))):
            z0 = int(rng.integers(0, 1 << n))
            k0 = complex(rng.normal(), rng.normal())
            for e, se in elements:
                if (z0 ^ e) in members:
                    consistent = False
                members[z0 ^ e] = se * k0
        if not consistent:
            continue
        planted[xg] = gens
        for zt, kt in members.items():
            y = bin(xg & zt).count('1')
            xs.append(xg)
            zs.append(zt)
            cs.append(kt * [1, 1j, -1, -1j][y % 4])  # Pauli coefficient whose phased value is kt
    x = numpy.array(xs, dtype=numpy.uint64)
    z = numpy.array(zs, dtype=numpy.uint64)
    c = numpy.array(cs, dtype=numpy.complex128)
    group_x, mask1, mask2, flags = support.group_conditions(x, z, c)
    for g, xg in enumerate(group_x.tolist()):
        want = min(len(planted[xg]), 2)
        assert (flags[g] & 1) + ((flags[g] >> 2) & 1) >= want
    xr, zr, cr = support.reduce_terms(x, z, c, group_x, mask1, mask2, flags)
    j = numpy.arange(1 << n, dtype=numpy.uint64)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    k, kr = support._phased(x, z, c), support._phased(xr, zr, cr)
    full = numpy.zeros(1 << n, dtype=numpy.complex128)
    hinted = numpy.zeros(1 << n, dtype=numpy.complex128)
    for g, xg in enumerate(group_x):
        need = numpy.ones(1 << n, dtype=bool)
        if flags[g] & 1:
            need &= po.bit_parity(j & mask1[g]) == ((flags[g] >> 1) & 1)
        if flags[g] & 4:
            need &= po.bit_parity(j & mask2[g]) == ((flags[g] >> 3) & 1)
        gathered = v[(j ^ xg).astype(numpy.int64)]
        full += _coefficient_function(x, z, k, xg, n) * gathered
        hinted += numpy.where(need, _coefficient_function(xr, zr, kr, xg, n), 0.0) * gathered
    assert numpy.max(numpy.abs(hinted - full)) <= 1e-13 * numpy.max(numpy.abs(full))
    assert len(cr) < len(c)
