"""Host-side generators of the benchmark / test Hamiltonians as packed Pauli-mask tables.

The reference builds these with its symbolic layer (hamiltonians/hubbard.py:19-179,
testing/testing_utils.py:109-182 `random_interaction_operator`, then
transforms/opconversions/jordan_wigner.py).  That layer is out of scope and absent on the
GPU box, so the same operators are generated here directly in mask form with a
vectorised Jordan-Wigner expansion: every ladder operator is two Pauli strings,
    a_p  = 1/2 (X_p + i Y_p) Z_0..Z_{p-1},   a_p^dagger = 1/2 (X_p - i Y_p) Z_0..Z_{p-1},
a product of k ladder operators expands into 2^k strings multiplied with
    P(x1,z1) P(x2,z2) = i^{|x1&z1| + |x2&z2| + 2|z1&x2| - |x3&z3|} P(x1^x2, z1^z2),
and equal strings are summed.  tests/test_models.py checks term sets and coefficients
against golden vectors produced by the reference (tests/golden/).
"""
import itertools

import numpy

EQ_TOLERANCE = 1e-8

_M1 = numpy.uint64(0x5555555555555555)
_M2 = numpy.uint64(0x3333333333333333)
_M4 = numpy.uint64(0x0F0F0F0F0F0F0F0F)
_H01 = numpy.uint64(0x0101010101010101)


def popcount64(a):
    a = numpy.asarray(a, dtype=numpy.uint64)
    a = a - ((a >> numpy.uint64(1)) & _M1)
    a = (a & _M2) + ((a >> numpy.uint64(2)) & _M2)
    a = (a + (a >> numpy.uint64(4))) & _M4
    return ((a * _H01) >> numpy.uint64(56)).astype(numpy.int64)


class MaskOperator:
    """A Pauli sum held as arrays: qubit-space masks (bit q = qubit q) and coefficients.

    Quacks like a QubitOperator for the linalg functions: `.terms` builds the reference's
    dict lazily (needed only by the CPU oracle in tests); the compiler takes the arrays.
    """

    def __init__(self, n_qubits, xq, zq, coeff):
        self.n_qubits = int(n_qubits)
        self.xq = numpy.asarray(xq, dtype=numpy.uint64)
        self.zq = numpy.asarray(zq, dtype=numpy.uint64)
        self.coeff = numpy.asarray(coeff, dtype=numpy.complex128)
        self._terms = None

    def __len__(self):
        return len(self.coeff)

    def count_qubits(self):
        both = int(numpy.bitwise_or.reduce(self.xq | self.zq)) if len(self.coeff) else 0
        return both.bit_length()

    def index_masks(self, n_qubits=None):
        """(x_mask, z_mask) with qubit q moved to index bit n-1-q
        (linalg/linear_qubit_operator.py:126)."""
        n = self.n_qubits if n_qubits is None else n_qubits
        x = numpy.zeros_like(self.xq)
        z = numpy.zeros_like(self.zq)
        for q in range(n):
            sh_in, sh_out = numpy.uint64(q), numpy.uint64(n - 1 - q)
            one = numpy.uint64(1)
            x |= ((self.xq >> sh_in) & one) << sh_out
            z |= ((self.zq >> sh_in) & one) << sh_out
        return x, z

    @property
    def terms(self):
        if self._terms is None:
            terms = {}
            real = bool(numpy.all(self.coeff.imag == 0.0))
            for xq, zq, c in zip(self.xq.tolist(), self.zq.tolist(), self.coeff.tolist()):
                key = []
                both = xq | zq
                q = 0
                while both >> q:
                    if (both >> q) & 1:
                        xb, zb = (xq >> q) & 1, (zq >> q) & 1
                        key.append((q, 'Y' if xb and zb else ('X' if xb else 'Z')))
                    q += 1
                terms[tuple(key)] = c.real if real else c
            self._terms = terms
        return self._terms

    def n_distinct_x(self):
        return len(numpy.unique(self.xq))


def _ladder_strings(mode, raising):
    """The two (x, z, coefficient) strings of a JW ladder operator."""
    bit = 1 << mode
    low = bit - 1
    # X_p Z_<p  and  Y_p Z_<p = P(x=bit, z=low|bit)
    return ((bit, low, 0.5), (bit, low | bit, -0.5j if raising else 0.5j))


_I_POW = numpy.array([1, 1j, -1, -1j], dtype=numpy.complex128)


def _mul_strings(x1, z1, c1, x2, z2, c2):
    x3, z3 = x1 ^ x2, z1 ^ z2
    k = popcount64(x1 & z1) + popcount64(x2 & z2) + 2 * popcount64(z1 & x2) - popcount64(x3 & z3)
    return x3, z3, c1 * c2 * _I_POW[k % 4]


def jw_tables(n_qubits):
    """Per mode p the two Pauli strings of its ladder operators under Jordan-Wigner:
    A = X_p Z_<p and B = Y_p Z_<p, with a_p^(dagger) = (A -/+ ... ) i.e. 1/2 (A -+ i B)."""
    one = numpy.uint64(1)
    p = numpy.arange(n_qubits, dtype=numpy.uint64)
    bit = one << p
    low = bit - one
    return bit, low, bit, low | bit


def _bk_sets(index, n_qubits):
    """Fenwick-tree update / occupation / parity sets of the Bravyi-Kitaev transform
    (transforms/opconversions/bravyi_kitaev.py:54-99), as bit masks."""
    update = 0
    i = index + 1
    i += i & -i
    while i <= n_qubits:
        update |= 1 << (i - 1)
        i += i & -i
    occupation = 1 << index
    i = index + 1
    parent = i & (i - 1)
    i -= 1
    while i != parent:
        occupation |= 1 << (i - 1)
        i &= i - 1
    parity = 0
    i = index
    while i > 0:
        parity |= 1 << (i - 1)
        i &= i - 1
    return update, occupation, parity


def bk_tables(n_qubits):
    """Bravyi-Kitaev ladder strings (bravyi_kitaev.py:178-212): A = X_{U+j} Z_P,
    B = Y_j X_U Z_{(P^O)-j}; a_j^dagger = 1/2 (A - i B), a_j = 1/2 (A + i B)."""
    xa, za, xb, zb = [], [], [], []
    for j in range(n_qubits):
        update, occupation, parity = _bk_sets(j, n_qubits)
        bit = 1 << j
        xa.append(update | bit)
        za.append(parity)
        xb.append(update | bit)
        zb.append(((parity ^ occupation) & ~bit) | bit)
    as_u64 = lambda v: numpy.array(v, dtype=numpy.uint64)  # noqa: E731
    return as_u64(xa), as_u64(za), as_u64(xb), as_u64(zb)


def ladder_products(modes, kinds, coeffs, tables):
    """Qubit image of sum_t coeffs[t] * prod_j ladder(modes[t, j], kinds[j]) for an encoding
    given by its per-mode string tables (jw_tables / bk_tables).

    modes: int array (T, k); kinds: length-k tuple of 1 (raising) / 0 (lowering).
    Returns the un-collected (x, z, coefficient) arrays of all 2^k expansions.
    """
    xa, za, xb, zb = tables
    modes = numpy.asarray(modes, dtype=numpy.int64).reshape(len(coeffs), -1)
    coeffs = numpy.asarray(coeffs, dtype=numpy.complex128)
    k = modes.shape[1]
    xs, zs, cs = [], [], []
    for choice in itertools.product((0, 1), repeat=k):
        x = numpy.zeros(len(coeffs), dtype=numpy.uint64)
        z = numpy.zeros(len(coeffs), dtype=numpy.uint64)
        c = coeffs.copy()
        for j in range(k):
            m = modes[:, j]
            if choice[j] == 0:
                xj, zj, cj = xa[m], za[m], 0.5
            else:
                xj, zj, cj = xb[m], zb[m], (-0.5j if kinds[j] else 0.5j)
            x, z, c = _mul_strings(x, z, c, xj, zj, numpy.complex128(cj))
        xs.append(x)
        zs.append(z)
        cs.append(c)
    return numpy.concatenate(xs), numpy.concatenate(zs), numpy.concatenate(cs)


def jordan_wigner_products(modes, kinds, coeffs, n_qubits):
    """JW image of sum_t coeffs[t] * prod_j ladder(modes[t, j], kinds[j])."""
    return ladder_products(modes, kinds, coeffs, jw_tables(n_qubits))


def _collect(n_qubits, pieces, tol=EQ_TOLERANCE):
    """Sums equal strings, drops |c| < tol (the reference's += rule), sorts by (x, z)."""
    x = numpy.concatenate([p[0] for p in pieces])
    z = numpy.concatenate([p[1] for p in pieces])
    c = numpy.concatenate([p[2] for p in pieces])
    if n_qubits > 64:
        raise ValueError('generators support at most 64 qubits')
    if n_qubits <= 32:
        key = (x << numpy.uint64(32)) | z
        uniq, inverse = numpy.unique(key, return_inverse=True)
        ux, uz = uniq >> numpy.uint64(32), uniq & numpy.uint64(0xFFFFFFFF)
    else:  # same (x, z) lexicographic order, two-column form
        pairs, inverse = numpy.unique(numpy.stack([x, z], axis=1), axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        ux, uz = pairs[:, 0], pairs[:, 1]
    re = numpy.bincount(inverse, weights=c.real, minlength=len(ux))
    im = numpy.bincount(inverse, weights=c.imag, minlength=len(ux))
    total = re + 1j * im
    keep = numpy.abs(total) >= tol
    return MaskOperator(n_qubits, ux[keep], uz[keep], total[keep])


def _build(n_qubits, specs, tables, device=False, tol=EQ_TOLERANCE):
    """MaskOperator of a list of ladder-product pieces [(modes (T, k), kinds, coeffs)].

    device=False: numpy expansion + merge on the host (ladder_products, _collect).
    device=True: the same expansion, merge and 1e-8 rule on the GPU (csrc/termgen.cu, SURVEY.md
    section 8f row N3): one thread per (term, choice), stable radix sort, sequential run sums in
    the host builder's order -- the two give identical arrays."""
    if not device:
        return _collect(n_qubits, [ladder_products(m, kinds, c, tables) for m, kinds, c in specs], tol)
    import ctypes
    from openfermion_b200 import _lib
    _lib.require_device()
    if n_qubits > 64:
        raise ValueError('generators support at most 64 qubits')
    xa, za, xb, zb = (numpy.ascontiguousarray(t, dtype=numpy.uint64) for t in tables)
    gen = ctypes.c_void_p()
    _lib.call('ofb_termgen_create', ctypes.c_int(n_qubits), _lib.ptr(xa), _lib.ptr(za), _lib.ptr(xb),
              _lib.ptr(zb), ctypes.c_int(0), ctypes.byref(gen))
    try:
        for modes, kinds, coeffs in specs:
            coeffs = numpy.ascontiguousarray(coeffs, dtype=numpy.complex128)
            modes = numpy.ascontiguousarray(numpy.asarray(modes, dtype=numpy.int32).reshape(len(coeffs), -1))
            k = len(kinds)
            mask = sum(1 << j for j, kind in enumerate(kinds) if kind)
            _lib.call('ofb_termgen_add', gen, ctypes.c_int(k), ctypes.c_uint32(mask),
                      ctypes.c_int64(len(coeffs)), _lib.ptr(modes) if k else None, _lib.ptr(coeffs))
        count = ctypes.c_int64()
        _lib.call('ofb_termgen_collect', gen, ctypes.c_double(tol), ctypes.byref(count))
        x = numpy.empty(count.value, dtype=numpy.uint64)
        z = numpy.empty(count.value, dtype=numpy.uint64)
        c = numpy.empty(count.value, dtype=numpy.complex128)
        _lib.call('ofb_termgen_fetch', gen, _lib.ptr(x), _lib.ptr(z), _lib.ptr(c))
    finally:
        _lib.call('ofb_termgen_destroy', gen)
    return MaskOperator(n_qubits, x, z, c)


# ----------------------------------------------------------------------------
# Fermi-Hubbard (hamiltonians/hubbard.py:128-179, spinful, up = 2*site, down = 2*site + 1)
# ----------------------------------------------------------------------------
def hubbard_fermion_terms(x_dimension, y_dimension, tunneling, coulomb, chemical_potential=0.0,
                          magnetic_field=0.0, periodic=True):
    """[(modes tuple, kinds tuple, coefficient)] of the spinful Hubbard model."""
    n_sites = x_dimension * y_dimension
    out = []

    def right(site):
        if x_dimension == 1:
            return None
        if (site + 1) % x_dimension == 0:
            return site + 1 - x_dimension if periodic else None
        return site + 1

    def bottom(site):
        if y_dimension == 1:
            return None
        if site + x_dimension + 1 > n_sites:
            return site + x_dimension - n_sites if periodic else None
        return site + x_dimension

    for site in range(n_sites):
        r, b = right(site), bottom(site)
        if x_dimension == 2 and periodic and site % 2 == 1:
            r = None
        if y_dimension == 2 and periodic and site >= x_dimension:
            b = None
        for other in (r, b):
            if other is None:
                continue
            for spin in (0, 1):
                i, j = 2 * site + spin, 2 * other + spin
                out.append(((i, j), (1, 0), -tunneling))
                out.append(((j, i), (1, 0), -tunneling))
        up, down = 2 * site, 2 * site + 1
        out.append(((up, up, down, down), (1, 0, 1, 0), coulomb))
        if chemical_potential or magnetic_field:
            out.append(((up, up), (1, 0), -chemical_potential - magnetic_field))
            out.append(((down, down), (1, 0), -chemical_potential + magnetic_field))
    return out


def _tables_for(encoding, n_qubits):
    if encoding in ('jw', 'jordan_wigner'):
        return jw_tables(n_qubits)
    if encoding in ('bk', 'bravyi_kitaev'):
        return bk_tables(n_qubits)
    raise ValueError('encoding must be jw or bk')


def _jw_of_term_list(term_list, n_qubits, encoding='jw', device=False):
    tables = _tables_for(encoding, n_qubits)
    by_kind = {}
    for modes, kinds, coefficient in term_list:
        by_kind.setdefault(kinds, ([], []))
        by_kind[kinds][0].append(modes)
        by_kind[kinds][1].append(coefficient)
    specs = [(numpy.asarray(m, dtype=numpy.int64).reshape(len(c), -1), kinds, c)
             for kinds, (m, c) in by_kind.items()]
    return _build(n_qubits, specs, tables, device)


def hubbard_jw(x_dimension, y_dimension, tunneling=1.0, coulomb=4.0, chemical_potential=0.0,
               magnetic_field=0.0, periodic=True, encoding='jw', device=False):
    """jordan_wigner(fermi_hubbard(x, y, t, U)) as a MaskOperator (2*x*y qubits);
    encoding='bk' gives bravyi_kitaev(...) instead; device=True expands and merges the strings on
    the GPU (csrc/termgen.cu)."""
    n_qubits = 2 * x_dimension * y_dimension
    terms = hubbard_fermion_terms(x_dimension, y_dimension, tunneling, coulomb, chemical_potential,
                                  magnetic_field, periodic)
    return _jw_of_term_list(terms, n_qubits, encoding, device)


# ----------------------------------------------------------------------------
# synthetic molecular-type Hamiltonian (testing/testing_utils.py:109-182)
# ----------------------------------------------------------------------------
def random_interaction_tensors(n_orbitals, seed=None, dyadic_bits=None):
    """(constant, one_body, two_body) spin-orbital tensors of
    random_interaction_operator(n_orbitals, expand_spin=True, real=True, seed) with the same
    random stream and symmetrisation order.  dyadic_bits=k rounds every drawn number to a
    multiple of 2^-k (exactly summable coefficients, SURVEY.md section 8c finding 3)."""
    if seed is not None:
        numpy.random.seed(seed)

    def rnd(value):
        if dyadic_bits is None:
            return value
        return numpy.round(value * 2.0**dyadic_bits) / 2.0**dyadic_bits

    constant = rnd(numpy.random.randn())
    # random_hermitian_matrix(n, real=True): randn matrix symmetrised (testing_utils.py:48-57)
    mat = rnd(numpy.random.randn(n_orbitals, n_orbitals))
    one_body = mat + mat.T
    n = n_orbitals
    two_body = numpy.zeros((n, n, n, n))
    draws = rnd(numpy.random.randn(n**4))
    for idx, (p, q, r, s) in enumerate(itertools.product(range(n), repeat=4)):
        coeff = draws[idx]
        two_body[p, q, r, s] = coeff
        two_body[q, p, s, r] = coeff
        two_body[s, r, q, p] = coeff
        two_body[r, s, p, q] = coeff
        two_body[r, q, p, s] = coeff
        two_body[p, s, r, q] = coeff
        two_body[s, p, q, r] = coeff
        two_body[q, r, s, p] = coeff
    one_spin = numpy.kron(one_body, numpy.eye(2))
    m = 2 * n
    two_spin = numpy.zeros((m, m, m, m))
    a = numpy.arange(n)
    P, Q, R, S = numpy.meshgrid(a, a, a, a, indexing='ij')
    two_spin[2 * P, 2 * Q + 1, 2 * R + 1, 2 * S] = two_body
    two_spin[2 * P + 1, 2 * Q, 2 * R, 2 * S + 1] = two_body
    two_spin[2 * P, 2 * Q, 2 * R, 2 * S] = two_body
    two_spin[2 * P + 1, 2 * Q + 1, 2 * R + 1, 2 * S + 1] = two_body
    return constant, one_spin, two_spin


def interaction_operator_jw(constant, one_body, two_body, encoding='jw', n_qubits=None, device=False):
    """JW of constant + sum h_pq p^ q + sum h_pqrs p^ q^ r s (InteractionOperator convention,
    ops/representations/interaction_operator.py) as a MaskOperator; device=True expands and merges
    the strings on the GPU (csrc/termgen.cu, same arrays)."""
    one_body = numpy.asarray(one_body)
    two_body = numpy.asarray(two_body)
    if n_qubits is None:
        n_qubits = one_body.shape[0]
    tables = _tables_for(encoding, n_qubits)
    specs = []
    if constant:  # the identity string: a product of zero ladder operators
        specs.append((numpy.zeros((1, 0), dtype=numpy.int64), (), numpy.array([constant], dtype=numpy.complex128)))
    idx1 = numpy.argwhere(one_body != 0)
    if len(idx1):
        specs.append((idx1, (1, 0), one_body[tuple(idx1.T)]))
    idx2 = numpy.argwhere(two_body != 0)
    if len(idx2):
        specs.append((idx2, (1, 1, 0, 0), two_body[tuple(idx2.T)]))
    return _build(n_qubits, specs, tables, device)


def molecular_jw(n_orbitals, seed=1234, dyadic_bits=None, encoding='jw', device=False):
    """The synthetic molecular-type qubit Hamiltonian of BASELINE.json (2*n_orbitals qubits):
    jordan_wigner(random_interaction_operator(n_orbitals, expand_spin=True, real=True, seed));
    encoding='bk' applies bravyi_kitaev to the same fermion operator; device=True generates the
    term table on the GPU (csrc/termgen.cu)."""
    return interaction_operator_jw(*random_interaction_tensors(n_orbitals, seed, dyadic_bits),
                                   encoding=encoding, device=device)


def random_pauli_sum(n_qubits, n_terms, seed=0, complex_coefficients=False):
    """Random dense Pauli strings (each qubit I/X/Y/Z uniformly), the generator of the
    reference's benchmark_linear_qubit_operator (testing/performance_benchmarks.py:130-175)."""
    rng = numpy.random.default_rng(seed)
    x = numpy.zeros(n_terms, dtype=numpy.uint64)
    z = numpy.zeros(n_terms, dtype=numpy.uint64)
    for q in range(n_qubits):
        kind = rng.integers(0, 4, size=n_terms)
        x |= (((kind == 1) | (kind == 2)).astype(numpy.uint64)) << numpy.uint64(q)
        z |= (((kind == 2) | (kind == 3)).astype(numpy.uint64)) << numpy.uint64(q)
    c = rng.normal(size=n_terms).astype(numpy.complex128)
    if complex_coefficients:
        c = c + 1j * rng.normal(size=n_terms)
    return _collect(n_qubits, [(x, z, c)], tol=0.0)


# ----------------------------------------------------------------------------
# molecular Hamiltonians from spatial integrals (chem/molecular_data.py:365-400, 1005-1041)
# ----------------------------------------------------------------------------
class InteractionTensor:
    """Minimal stand-in for the reference's InteractionOperator (a PolynomialTensor with
    keys (), (1, 0), (1, 1, 0, 0)); accepted by linalg.get_sparse_operator."""

    def __init__(self, constant, one_body_tensor, two_body_tensor):
        self.constant = constant
        self.one_body_tensor = numpy.asarray(one_body_tensor)
        self.two_body_tensor = numpy.asarray(two_body_tensor)
        self.n_qubits = self.one_body_tensor.shape[0]
        self.n_body_tensors = {(): constant, (1, 0): self.one_body_tensor,
                               (1, 1, 0, 0): self.two_body_tensor}


def spin_orbital_tensors(one_body_integrals, two_body_integrals):
    """Spatial -> spin-orbital coefficients, even = up, odd = down, entries below 1e-8
    truncated (chem/molecular_data.py:365-400)."""
    one_body_integrals = numpy.asarray(one_body_integrals)
    two_body_integrals = numpy.asarray(two_body_integrals)
    n = one_body_integrals.shape[0]
    m = 2 * n
    one = numpy.zeros((m, m))
    two = numpy.zeros((m, m, m, m))
    one[0::2, 0::2] = one_body_integrals
    one[1::2, 1::2] = one_body_integrals
    two[0::2, 1::2, 1::2, 0::2] = two_body_integrals
    two[1::2, 0::2, 0::2, 1::2] = two_body_integrals
    two[0::2, 0::2, 0::2, 0::2] = two_body_integrals
    two[1::2, 1::2, 1::2, 1::2] = two_body_integrals
    one[numpy.absolute(one) < EQ_TOLERANCE] = 0.0
    two[numpy.absolute(two) < EQ_TOLERANCE] = 0.0
    return one, two


def molecular_hamiltonian(nuclear_repulsion, one_body_integrals, two_body_integrals):
    """MolecularData.get_molecular_hamiltonian() without an active space
    (chem/molecular_data.py:1021-1039)."""
    one, two = spin_orbital_tensors(one_body_integrals, two_body_integrals)
    return InteractionTensor(nuclear_repulsion, one, 0.5 * two)


class DiagonalCoulomb:
    """Minimal stand-in for the reference's DiagonalCoulombHamiltonian
    (ops/representations/diagonal_coulomb_hamiltonian.py): sum T_pq p^ q + sum V_pq n_p n_q + c."""

    def __init__(self, one_body, two_body, constant=0.0):
        self.one_body = numpy.asarray(one_body)
        self.two_body = numpy.asarray(two_body)
        self.constant = constant
