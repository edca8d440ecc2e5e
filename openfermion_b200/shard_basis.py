"""Symmetry-adapted sharding: a GF(2)-linear relabelling of the basis states that puts
conserved parities on the shard bits.

The sharded H|psi> (distributed.py) splits the state by the top k index bits; an x-group
whose mask touches those bits needs a whole peer shard.  Which k functions of the basis
index are "the top bits" is a free choice: for any invertible binary matrix B the
relabelling c' = B c (arithmetic mod 2 on the index bits) maps a Pauli string onto a Pauli
string,

    P = k * sum_c (-1)^{z.c} |c ^ x><c|     ->     P' = k * sum_c' (-1)^{z'.c'} |c' ^ x'><c'|
    x' = B x,     z' = B^{-T} z,     k = coeff * i^{|x & z|}  unchanged

(the reference's convention for x, z and the i^y phase: linalg/linear_qubit_operator.py:
122-145), so the same kernels run on the relabelled state psi'[c'] = psi[B^{-1} c'] with
transformed mask tables.  If a row m of B is orthogonal to every x-mask of the Hamiltonian
(m.x = 0 mod 2 for all groups: the parity of the qubits in m is conserved -- the up- and
down-spin particle-number parities of every spin-conserving fermionic Hamiltonian under
Jordan-Wigner, the total parity under any encoding) and that row is made a shard bit, NO
group crosses it.  The 4x4 Hubbard model at 32 qubits then needs no exchange at all on
2 and 4 GPUs and one peer shard instead of four on 8 GPUs.

Host-side only (numpy and Python integers); the reference has no counterpart (it never
shards a state).  The relabelling is a permutation of the amplitudes, so results agree with
the natural order to the last bit after `ShardBasis.to_natural`.
"""
import numpy

from openfermion_b200.hamiltonians import popcount64  # portable (numpy.bitwise_count needs numpy >= 2)

_ONE = numpy.uint64(1)


def _parity64(a):
    return (popcount64(a) & 1).astype(numpy.uint64)


def _reduce_gf2(vector, pivots):
    """Reduce `vector` by an echelon set {pivot_bit: row} (each row's top set bit is its key)."""
    while vector:
        top = vector.bit_length() - 1
        row = pivots.get(top)
        if row is None:
            return vector
        vector ^= row
    return 0


def gf2_nullspace(masks, n_bits):
    """Basis of {m : popc(m & x) even for every x in masks} as Python ints (n_bits wide)."""
    # row-reduce the x-masks to reduced echelon form, then read off one vector per free column
    pivots = {}
    for x in sorted({int(v) for v in masks}, reverse=True):
        x = _reduce_gf2(x, pivots)
        if x:
            pivots[x.bit_length() - 1] = x
    # back-substitute so that every pivot column appears in exactly one row
    for top in sorted(pivots):
        for other in pivots:
            if other != top and (pivots[other] >> top) & 1:
                pivots[other] ^= pivots[top]
    basis = []
    for free in range(n_bits - 1, -1, -1):
        if free in pivots:
            continue
        m = 1 << free
        for top, row in pivots.items():
            if (row >> free) & 1:
                m |= 1 << top
        basis.append(m)
    return basis


def _distinct_outside(groups, low_mask):
    """Number of distinct parts of the x-masks outside `low_mask`: how many different input
    blocks (own block included) a block of the index space spanned by `low_mask` gathers from."""
    return len(numpy.unique(groups & ~numpy.uint64(low_mask)))


def _best_subset(groups, pool, size, starts):
    """Subset of `pool` (bit positions) of the given size minimising _distinct_outside, by
    steepest-descent swaps from a few deterministic starting sets."""
    if size >= len(pool):
        return sorted(pool)
    best = None
    for start in starts:
        cur = sorted(start)[:size]
        cur += [b for b in pool if b not in cur][:size - len(cur)]
        mask = sum(1 << b for b in cur)
        cost = _distinct_outside(groups, mask)
        improved = True
        while improved:
            improved = False
            move = None
            for out_bit in cur:
                base = mask & ~(1 << out_bit)
                for in_bit in pool:
                    if (mask >> in_bit) & 1:
                        continue
                    c = _distinct_outside(groups, base | (1 << in_bit))
                    if c < cost:
                        cost, move = c, (out_bit, in_bit)
            if move is not None:
                cur[cur.index(move[0])] = move[1]
                mask = sum(1 << b for b in cur)
                improved = True
        if best is None or cost < best[0]:
            best = (cost, sorted(cur))
    return best[1]


# Levels of the memory hierarchy the apply kernel (csrc/apply.cu) re-uses input blocks in:
# a CTA tile of 2^10 amplitudes (L1) and the span of about 2^19 amplitudes that the CTAs in
# flight keep alive in L2 (profiles/README.md: for the 30-qubit Hubbard model the number of
# distinct x >> 10 values, 45 of 61 groups, matches the measured L1 miss share and the number
# of distinct x >> 18..19 values, 27..31, the measured 30 DRAM read streams).
LOCALITY_LEVELS = (10, 19)
LOCALITY_MAX_GROUPS = 4096


def locality_order(x_masks, n_bits, levels=LOCALITY_LEVELS):
    """Order of the index bits (a list of natural bit positions, new bit 0 first) that
    minimises the number of distinct x-mask parts above each level: first the best
    `levels[-1]`-bit set, then the best `levels[0]` bits inside it; bits keep their natural
    relative order inside a level.  The natural order is returned when less than 10 % is gained
    or the Hamiltonian has too many groups for locality to matter (dense molecular masks)."""
    groups = numpy.unique(numpy.asarray(x_masks, dtype=numpy.uint64))
    natural = list(range(n_bits))
    if len(groups) < 2 or len(groups) > LOCALITY_MAX_GROUPS:
        return natural
    sizes = [lv for lv in sorted(levels) if lv < n_bits]
    if not sizes:
        return natural
    # greedy forward order as a second starting point
    greedy, mask = [], 0
    for _ in range(n_bits):
        pick = min((b for b in natural if not (mask >> b) & 1),
                   key=lambda b: (_distinct_outside(groups, mask | (1 << b)), b))
        greedy.append(pick)
        mask |= 1 << pick
    pool = natural
    nested = []
    for size in reversed(sizes):
        chosen = _best_subset(groups, pool, size, [pool, [b for b in greedy if b in pool]])
        nested.append(chosen)
        pool = chosen
    # nested[-1] is the innermost (smallest) set
    order = list(nested[-1])
    for outer in reversed(nested[:-1]):
        order += [b for b in outer if b not in order]
    order += [b for b in natural if b not in order]

    def costs(o):
        return tuple(_distinct_outside(groups, sum(1 << b for b in o[:size])) for size in sizes)
    # re-order only for a clear gain (at least 10 % fewer gathered blocks over the levels)
    return order if sum(costs(order)) <= 0.9 * sum(costs(natural)) else natural


class ShardBasis:
    """c' = B c over GF(2).  `rows[b]` is the mask whose parity with c gives bit b of c';
    `cols[b]` is the natural index of the relabelled basis state 1 << b (column b of B^-1)."""

    def __init__(self, n_qubits, rows):
        if len(rows) != n_qubits:
            raise ValueError('need one row per index bit')
        self.n_qubits = n_qubits
        self.rows = [int(r) for r in rows]
        self.cols = self._invert(self.rows, n_qubits)
        self.inv_rows = [0] * n_qubits  # row i of B^-1: bit b = entry (i, b)
        for b, col in enumerate(self.cols):
            for i in range(n_qubits):
                if (col >> i) & 1:
                    self.inv_rows[i] |= 1 << b
        self.identity = all(r == 1 << b for b, r in enumerate(self.rows))

    @staticmethod
    def _invert(rows, n):
        # Gauss-Jordan on [B | I] with rows held as integers: bit b of work[i] = B[i][b]
        work = list(rows)
        inv = [1 << i for i in range(n)]
        for col in range(n):
            pivot = next((i for i in range(col, n) if (work[i] >> col) & 1), None)
            if pivot is None:
                raise ValueError('shard basis is not invertible')
            work[col], work[pivot] = work[pivot], work[col]
            inv[col], inv[pivot] = inv[pivot], inv[col]
            for i in range(n):
                if i != col and (work[i] >> col) & 1:
                    work[i] ^= work[col]
                    inv[i] ^= inv[col]
        # inv[i] holds row i of B^-1 (bit j = entry (i, j)); column b of B^-1 as a mask over i
        cols = []
        for b in range(n):
            c = 0
            for i in range(n):
                if (inv[i] >> b) & 1:
                    c |= 1 << i
            cols.append(c)
        return cols

    @classmethod
    def natural(cls, n_qubits):
        return cls(n_qubits, [1 << b for b in range(n_qubits)])

    @classmethod
    def for_sharding(cls, n_qubits, shard_bits, x_masks, local_order='natural'):
        """Choose the `shard_bits` top rows to minimise the number of distinct non-zero crossing
        patterns (= peer shards pulled per matvec), greedily: conserved parities (null space of
        the x-masks) first, then single qubits, highest first.  Ties go to fewer crossing groups,
        then to the earlier candidate.  The remaining rows keep the other index bits in their
        natural order, so a shard's local layout is the natural one -- unless
        local_order='locality', which also permutes the local bits so that as many x-groups as
        possible gather from the same cache-resident blocks (`locality_order`)."""
        n = n_qubits
        groups = numpy.unique(numpy.asarray(x_masks, dtype=numpy.uint64))
        if local_order not in ('natural', 'locality'):
            raise ValueError("local_order must be 'natural' or 'locality'")
        if n == 0 or (shard_bits == 0 and local_order == 'natural'):
            return cls.natural(n)
        candidates = gf2_nullspace(groups.tolist(), n) + [1 << b for b in range(n - 1, -1, -1)]
        chosen, echelon = [], {}
        pattern = numpy.zeros(len(groups), dtype=numpy.uint64)  # crossing pattern per group so far
        for _ in range(shard_bits):
            best = None
            for order, m in enumerate(candidates):
                if not _reduce_gf2(m, echelon):
                    continue  # dependent on the rows already chosen
                bit = _parity64(groups & numpy.uint64(m))
                trial = (pattern << _ONE) | bit
                distinct = numpy.unique(trial)
                cost = (int(numpy.count_nonzero(distinct)), int(numpy.count_nonzero(trial)), order)
                if best is None or cost < best[0]:
                    best = (cost, m, trial)
                if cost[0] == 0:
                    break
            if best is None:
                raise ValueError('cannot find {} independent shard rows'.format(shard_bits))
            _, m, pattern = best
            chosen.append(m)
            r = _reduce_gf2(m, echelon)
            echelon[r.bit_length() - 1] = r
        # pivots = the index bits the shard rows replace; every other bit keeps its place
        pivots = sorted(echelon, reverse=True)
        kept = [b for b in range(n - 1, -1, -1) if b not in echelon]
        rows_top_down = chosen + [1 << b for b in kept]
        assert len(pivots) == shard_bits and len(rows_top_down) == n
        basis = cls(n, rows_top_down[::-1])
        if local_order == 'locality':
            # order the local bits by the x-masks as the kernels will see them
            local_bits = n - shard_bits
            local_x = basis.to_basis_index(groups) & numpy.uint64((1 << local_bits) - 1)
            order = locality_order(local_x, local_bits)
            rows = [basis.rows[b] for b in order] + basis.rows[local_bits:]
            basis = cls(n, rows)
        return basis

    # -- indices ---------------------------------------------------------------------------
    def _apply(self, table, index):
        index = numpy.asarray(index, dtype=numpy.uint64)
        out = numpy.zeros_like(index)
        for b, m in enumerate(table):
            out |= _parity64(index & numpy.uint64(m)) << numpy.uint64(b)
        return out

    def to_basis_index(self, c):
        """c' = B c for an array of natural indices."""
        return self._apply(self.rows, c)

    def to_natural_index(self, c_prime):
        """c = B^-1 c' : bit i of c is the parity of c' with row i of B^-1."""
        return self._apply(self.inv_rows, c_prime)

    def shard_natural_indices(self, rank, local_bits):
        """Natural index of every amplitude of rank `rank`'s shard, in shard order."""
        local = numpy.arange(1 << local_bits, dtype=numpy.uint64)
        return self.to_natural_index(local | numpy.uint64(rank << local_bits))

    def from_natural(self, state):
        """psi'[c'] = psi[B^-1 c'] (full-length host vector)."""
        state = numpy.asarray(state)
        if self.identity:
            return state.copy()
        index = self.to_natural_index(numpy.arange(len(state), dtype=numpy.uint64))
        return state[index.astype(numpy.int64)]

    def to_natural(self, state_prime):
        state_prime = numpy.asarray(state_prime)
        if self.identity:
            return state_prime.copy()
        index = self.to_natural_index(numpy.arange(len(state_prime), dtype=numpy.uint64))
        out = numpy.empty_like(state_prime)
        out[index.astype(numpy.int64)] = state_prime
        return out

    # -- mask tables -------------------------------------------------------------------------
    def transform_terms(self, x, z, coeff):
        """(x, z, coeff) of Pauli strings in the natural basis -> the same operator's strings in
        this basis.  x' = B x, z' = B^-T z; the library derives the phase i^{|x & z|} from the
        masks it is given, so the coefficient absorbs i^{|x & z| - |x' & z'|} (a sign: the
        parity of |x & z| is invariant)."""
        x = numpy.asarray(x, dtype=numpy.uint64)
        z = numpy.asarray(z, dtype=numpy.uint64)
        coeff = numpy.asarray(coeff, dtype=numpy.complex128)
        if self.identity:
            return x.copy(), z.copy(), coeff.copy()
        xp = self._apply(self.rows, x)
        zp = self._apply(self.cols, z)
        y = popcount64(x & z)
        yp = popcount64(xp & zp)
        delta = (y - yp) % 4
        if numpy.any(delta & 1):
            raise AssertionError('relabelling changed the parity of a Y count')
        return xp, zp, numpy.where(delta == 2, -coeff, coeff)

    def describe(self, shard_bits):
        """Human-readable list of the shard bits, most significant first."""
        if self.identity:
            return 'top %d qubits' % shard_bits
        n = self.n_qubits
        parts = []
        for b in range(n - 1, n - 1 - shard_bits, -1):
            m = self.rows[b]
            qubits = [n - 1 - i for i in range(n - 1, -1, -1) if (m >> i) & 1]
            parts.append('qubit %d' % qubits[0] if len(qubits) == 1
                         else 'parity of %d qubits (%d..%d)' % (len(qubits), qubits[0], qubits[-1]))
        return ', '.join(parts)
