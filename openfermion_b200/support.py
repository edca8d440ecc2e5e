"""Support hints for the matrix-free kernels: where an x-group's coefficient vanishes.

In the gather form  out[j] = sum_g S_g(j) in[j ^ x_g],  S_g(j) = sum_{t in g} k_t (-1)^{|j & z_t|}
(k_t = coeff_t (-i)^{|x & z_t|}), the coefficient S_g(j) of a symmetric Hamiltonian is zero on
most of the index space: a hopping group X_a X_b + Y_a Y_b (both with the same Z string) only
connects basis states whose bits a and b differ, a double excitation only states with the right
occupation pattern.  The reference never notices (it applies term by term,
linalg/linear_qubit_operator.py:107-148); a kernel that gathers in[j ^ x] for every j fetches
two to four times more amplitudes than contribute.

Algebra.  If the term set of a group is invariant under  z -> z ^ m,  k -> s k  (s = +-1), then
chi_m(j) S_g(j) = s S_g(j)  with  chi_m(j) = (-1)^{|j & m|},  so

    S_g(j) = 0   unless   parity(j & m) == (1 if s == -1 else 0).

`group_conditions` finds up to two independent such (m, s) per group directly from the term
table -- no knowledge of spin or particle number is needed, and encodings other than
Jordan-Wigner simply yield whatever symmetry their strings have.  The conditions go to the
library with `ofb_plan_set_support` (include/ofb200.h); the kernel then skips the gathers, and
whole warps skip groups whose condition does not depend on the lane or class bits.

On by default since round 2 (measured on B200: 28-qubit molecular H|psi> 2652 -> 1817 ms, 30-qubit
Hubbard 81 -> 56 ms); OFB_SUPPORT_HINTS=0 turns it off.
"""
import numpy

from openfermion_b200.hamiltonians import popcount64  # portable (numpy.bitwise_count needs numpy >= 2)

# relative tolerance on "equal coefficients": the strings of one fermionic excitation get their
# coefficients from the same integrals but possibly through differently ordered sums
COEFF_RTOL = 8 * numpy.finfo(numpy.float64).eps


def _phased(x, z, c):
    """Gather-form coefficients k_t = c_t (-i)^{|x & z_t|} (what the kernels sum)."""
    y = popcount64(numpy.asarray(x, dtype=numpy.uint64) & numpy.asarray(z, dtype=numpy.uint64))
    return numpy.asarray(c, dtype=numpy.complex128) * numpy.array([1, -1j, -1, 1j])[y % 4]


def _group_symmetries(zs, ks):
    """All (m, v) with S(j) = 0 unless parity(j & m) == v, for one group (python ints / complex)."""
    if len(zs) < 2:
        return []
    index = {}
    for t, zt in enumerate(zs):
        if zt in index:  # the same string twice: merge first (cannot happen for canonical operators)
            return []
        index[zt] = t
    scale = max(abs(k) for k in ks)
    if scale == 0.0:
        return []
    tol = COEFF_RTOL * scale
    found = []
    for t in range(1, len(zs)):
        m = zs[0] ^ zs[t]
        for s in (1.0, -1.0):
            if abs(ks[t] - s * ks[0]) > tol:
                continue
            ok = True
            for u, zu in enumerate(zs):
                w = index.get(zu ^ m)
                if w is None or abs(ks[w] - s * ks[u]) > tol:
                    ok = False
                    break
            if ok:
                found.append((m, 0 if s > 0 else 1))
                break
    return found


def _cost(m, low_bits=10):
    """Prefer masks that do not touch the lane / class bits (whole warps can skip the group) and,
    among those that do, masks that leave bit 0 alone (32-byte sectors hold two amplitudes)."""
    return (bin(m & ((1 << low_bits) - 1)).count('1'), m & 1, m)


def group_conditions(x, z, c):
    """Per x-group (ascending x = the plan's group order): up to two independent conditions.

    Returns (group_x u64[G], mask1 u64[G], mask2 u64[G], flags u8[G]) with
    flags bit0: condition 1 present, bit1: its parity value v1, bit2 / bit3 likewise for 2."""
    x = numpy.asarray(x, dtype=numpy.uint64)
    z = numpy.asarray(z, dtype=numpy.uint64)
    k = _phased(x, z, c)
    order = numpy.argsort(x, kind='stable')
    xs = x[order]
    group_x, starts = numpy.unique(xs, return_index=True)
    bounds = list(starts) + [len(xs)]
    n_groups = len(group_x)
    mask1 = numpy.zeros(n_groups, dtype=numpy.uint64)
    mask2 = numpy.zeros(n_groups, dtype=numpy.uint64)
    flags = numpy.zeros(n_groups, dtype=numpy.uint8)
    z_list = z[order].tolist()
    k_list = k[order].tolist()
    for g in range(n_groups):
        b, e = bounds[g], bounds[g + 1]
        if int(group_x[g]) == 0:
            continue  # the diagonal group: no gather to skip
        sym = sorted(_group_symmetries(z_list[b:e], k_list[b:e]), key=lambda mv: _cost(mv[0]))
        if not sym:
            continue
        (m1, v1) = sym[0]
        mask1[g] = m1
        flags[g] |= 1 | (v1 << 1)
        for m2, v2 in sym[1:]:
            if m2 != m1:  # two distinct non-zero vectors over GF(2) are independent
                mask2[g] = m2
                flags[g] |= 4 | (v2 << 3)
                break
    return group_x, mask1, mask2, flags


def reduce_terms(x, z, c, group_x, mask1, mask2, flags, class_shift=6, class_mask=15):
    """Term table that equals the original one ON THE SUPPORT of every group, with one term per
    symmetry orbit.

    On the support of a group chi_m(j) = s for each of its conditions, so the strings z and z ^ m
    have the same sign pattern up to the constant s: the orbit {z, z ^ m1, z ^ m2, z ^ m1 ^ m2}
    collapses to its representative with the coefficient  sum_u s_u k_u  (= 2 k or 4 k for an
    exactly symmetric group).  A kernel that zeroes the amplitudes outside the support -- the
    support-hint variant does -- can therefore sum a quarter of the terms of a mixed-spin double
    excitation on a quarter of the amplitudes.  The table is only valid together with the hints.

    Returns (x', z', c') with Pauli coefficients c' (the library re-derives the phase from the
    masks), groups in ascending x, representatives in their original relative order."""
    x = numpy.asarray(x, dtype=numpy.uint64)
    z = numpy.asarray(z, dtype=numpy.uint64)
    k = _phased(x, z, c)
    order = numpy.argsort(x, kind='stable')
    xs = x[order].tolist()
    zs = z[order].tolist()
    ks = k[order].tolist()
    group_of = {int(gx): g for g, gx in enumerate(numpy.asarray(group_x).tolist())}
    out_x, out_z, out_k = [], [], []
    start = 0
    total = len(xs)
    while start < total:
        end = start
        while end < total and xs[end] == xs[start]:
            end += 1
        g = group_of[xs[start]]
        gens = []
        if flags[g] & 1:
            gens.append((int(mask1[g]), -1.0 if (flags[g] >> 1) & 1 else 1.0))
        if flags[g] & 4:
            gens.append((int(mask2[g]), -1.0 if (flags[g] >> 3) & 1 else 1.0))
        if not gens:
            out_x += xs[start:end]
            out_z += zs[start:end]
            out_k += ks[start:end]
            start = end
            continue
        # the elements of the symmetry group generated by the conditions: (mask, sign)
        elements = [(0, 1.0)]
        for m, s in gens:
            elements += [(e ^ m, se * s) for e, se in elements]
        index = {zt: t for t, zt in enumerate(zs[start:end])}
        orbits, seen = [], set()
        for zt in zs[start:end]:
            if zt in seen:
                continue
            members = []
            for e, _ in elements:
                if (zt ^ e) not in index:
                    raise ValueError('term table is not invariant under its own support hints')
                members.append(zt ^ e)
                seen.add(zt ^ e)
            orbits.append(members)
        # Any member can represent its orbit.  The kernel pays 2 * 16 multiply-adds per distinct
        # "class" (the z bits CLASS_SHIFT..CLASS_SHIFT+3) among a group's terms, so pick the
        # representatives that share as few classes as possible (greedy cover).
        def cls(zv):
            return (zv >> class_shift) & class_mask
        chosen = [None] * len(orbits)
        while any(r is None for r in chosen):
            counts = {}
            for o, members in enumerate(orbits):
                if chosen[o] is None:
                    for w in {cls(zv) for zv in members}:
                        counts[w] = counts.get(w, 0) + 1
            best = min(counts, key=lambda w: (-counts[w], w))
            for o, members in enumerate(orbits):
                if chosen[o] is None:
                    hit = [zv for zv in members if cls(zv) == best]
                    if hit:
                        chosen[o] = min(hit)
        for rep in chosen:
            coeff = 0.0
            for e, se in elements:
                coeff += se * ks[start + index[rep ^ e]]
            out_x.append(xs[start])
            out_z.append(rep)
            out_k.append(coeff)
        start = end
    xr = numpy.array(out_x, dtype=numpy.uint64)
    zr = numpy.array(out_z, dtype=numpy.uint64)
    kr = numpy.array(out_k, dtype=numpy.complex128)
    y = popcount64(xr & zr)
    cr = kr * numpy.array([1, 1j, -1, -1j])[y % 4]  # undo the gather-form phase: c = k (+i)^y
    return xr, zr, cr


def conditions_implied_by_sector(group_x, mask1, mask2, flags, plus, minus):
    """Keep only the conditions that hold automatically inside a symmetry sector.

    The sector kernels (csrc/sector.cu) evaluate a group at j only if j and j ^ x are both
    members, i.e. satisfy |j & plus_k| - |j & minus_k| = value_k for every constraint k.  Both
    memberships together force  parity(j & x & (plus_k | minus_k)) = (|x & plus_k| - |x & minus_k|)/2
    mod 2, and with it every GF(2) combination over k.  A group condition inside that span never
    fails where the sector kernel evaluates the group, so the orbit-reduced table
    (`reduce_terms`) can be used there without any change to the kernel.  Returns new flags."""
    flags = numpy.array(flags, dtype=numpy.uint8)
    out = numpy.zeros_like(flags)
    plus = [int(p) for p in plus]
    minus = [int(m) for m in minus]
    for g, xg in enumerate(numpy.asarray(group_x).tolist()):
        if not flags[g]:
            continue
        implied = {}
        base = []
        for p, m in zip(plus, minus):
            a, b = bin(xg & p).count('1'), bin(xg & m).count('1')
            if (a - b) % 2:
                base = None  # no member j has a member partner j ^ x: the group never acts
                break
            base.append((xg & (p | m), ((a - b) // 2) % 2))
        if base is None:
            out[g] = flags[g]  # anything is "implied" on an empty set
            continue
        for subset in range(1, 1 << len(base)):
            m_acc, v_acc = 0, 0
            for k, (mk, vk) in enumerate(base):
                if (subset >> k) & 1:
                    m_acc ^= mk
                    v_acc ^= vk
            if m_acc:
                implied[m_acc] = v_acc
        if flags[g] & 1 and implied.get(int(mask1[g])) == ((flags[g] >> 1) & 1):
            out[g] |= flags[g] & 0b0011
        if flags[g] & 4 and implied.get(int(mask2[g])) == ((flags[g] >> 3) & 1):
            out[g] |= flags[g] & 0b1100
    return out


def needed_fraction(flags):
    """Share of (group, amplitude) gathers that remain: 1/2 per condition."""
    flags = numpy.asarray(flags)
    if len(flags) == 0:
        return 1.0
    per_group = 0.5 ** ((flags & 1) + ((flags >> 2) & 1)).astype(float)
    return float(per_group.mean())
