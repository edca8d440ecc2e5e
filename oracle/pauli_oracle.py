"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (numpy/scipy restatement of the reference hot path).

This module restates, on the CPU, the algorithms of the reference files

    /root/reference/src/openfermion/linalg/linear_qubit_operator.py
    /root/reference/src/openfermion/linalg/sparse_tools.py
    /root/reference/src/openfermion/linalg/davidson.py

so that the CUDA product path can be parity-checked on a box that has no copy
of the reference.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product package `openfermion_b200` never does.

Parity pinning: every function below is checked against (a) the known-answer
vectors of the reference's own tests (tests/test_oracle_golden.py restates
them) and (b) outputs of the real reference imported through oracle/ref_shim.py
in the build container; the generated vectors are committed under
tests/golden/ by oracle/make_golden.py.

Inputs are plain `terms` dicts in the reference's own encoding
(`QubitOperator.terms`, ops/operators/symbolic_operator.py:58-65):
    {((qubit, 'X'|'Y'|'Z'), ...): coefficient}      () is the identity
and for fermions (`FermionOperator.terms`)
    {((mode, 1|0), ...): coefficient}               1 = raising, 0 = lowering
"""
import functools
import multiprocessing
import warnings

import numpy
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

_PHASES = (1, 1j, -1, -1j)


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def count_qubits_terms(terms):
    """1 + highest qubit/mode index (reference utils/operator_utils.py:166-172)."""
    n = 0
    for term in terms:
        for factor in term:
            n = max(n, factor[0] + 1)
    return n


def term_masks(term, n_qubits):
    """(x_mask, z_mask, y_count) of one Pauli string.

    Follows linear_qubit_operator.py:122-134: qubit q is bit n-1-q.
    """
    x_mask = z_mask = y_count = 0
    for qubit, action in term:
        bit = 1 << (n_qubits - 1 - qubit)
        if action == 'X':
            x_mask ^= bit
        elif action == 'Y':
            x_mask ^= bit
            z_mask ^= bit
            y_count += 1
        else:
            z_mask ^= bit
    return x_mask, z_mask, y_count


def bit_parity(values):
    """Parity of popcount per uint64 (linear_qubit_operator.py:27-35)."""
    v = values.copy()
    for shift in (32, 16, 8, 4, 2, 1):
        v ^= v >> numpy.uint64(shift)
    return v & numpy.uint64(1)


# ----------------------------------------------------------------------------
# a2: LinearQubitOperator._matvec  (linear_qubit_operator.py:107-148)
# ----------------------------------------------------------------------------
def matvec(terms, n_qubits, x):
    """out = sum_t coeff_t * i^{y_t} * P_t x, scatter form, term order = dict order."""
    arr = numpy.asarray(x)
    vec = arr.reshape(-1)
    if vec.size != 1 << n_qubits:
        raise ValueError('dimension mismatch')
    out = numpy.zeros(vec.size, dtype=complex)
    indices = None
    for term, coefficient in terms.items():
        x_mask, z_mask, y_count = term_masks(term, n_qubits)
        amplitudes = coefficient * _PHASES[y_count % 4] * vec
        if (x_mask or z_mask) and indices is None:
            indices = numpy.arange(vec.size, dtype=numpy.uint64)
        if z_mask:
            signs = 1 - 2 * bit_parity(indices & numpy.uint64(z_mask)).astype(numpy.int8)
            amplitudes = signs * amplitudes
        if x_mask:
            out[indices ^ numpy.uint64(x_mask)] += amplitudes
        else:
            out += amplitudes
    return out.reshape(arr.shape)


def matvec_sampled(terms, n_qubits, x, out_indices, chunk=4096):
    """(H x)[j] for the sampled output indices j only, straight from the reference-order term
    dict (no grouping, no compiled tables): the scatter form of linear_qubit_operator.py:107-148
    read backwards,

        out[j] = sum_t coeff_t * i^{y_t} * (-1)^{popc((j ^ x_t) & z_t)} * x[j ^ x_t],

    accumulated in term order.  Cost O(len(terms) * len(out_indices)), independent of 2^n, so it
    checks a 28-qubit result in seconds.  `x` may be any indexable of length 2^n."""
    out_indices = numpy.asarray(out_indices, dtype=numpy.uint64)
    count = len(terms)
    xm = numpy.zeros(count, dtype=numpy.uint64)
    zm = numpy.zeros(count, dtype=numpy.uint64)
    ph = numpy.zeros(count, dtype=complex)
    for k, (term, coefficient) in enumerate(terms.items()):
        x_mask, z_mask, y_count = term_masks(term, n_qubits)
        xm[k], zm[k] = x_mask, z_mask
        ph[k] = coefficient * _PHASES[y_count % 4]
    if callable(x):  # amplitudes given as a function of the index (states too large for the host)
        lookup = lambda index: numpy.asarray(x(index.reshape(-1))).reshape(index.shape)
    else:
        vec = numpy.asarray(x).reshape(-1)
        lookup = lambda index: vec[index.astype(numpy.int64)]
    out = numpy.zeros(len(out_indices), dtype=complex)
    for lo in range(0, count, chunk):
        src = out_indices[:, None] ^ xm[None, lo:lo + chunk]          # the i with i ^ x_t == j
        signs = 1.0 - 2.0 * bit_parity(src & zm[None, lo:lo + chunk]).astype(numpy.float64)
        contrib = (signs * ph[None, lo:lo + chunk]) * lookup(src)
        for k in range(contrib.shape[1]):   # term order, like the reference's accumulation
            out += contrib[:, k]
    return out


def matmat(terms, n_qubits, X):
    """Column-by-column matvec (scipy LinearOperator._matmat default)."""
    X = numpy.asarray(X)
    return numpy.stack([matvec(terms, n_qubits, X[:, k]) for k in range(X.shape[1])], axis=1)


def split_terms(terms, groups):
    """Contiguous slices of the term dict (symbolic_operator.py:779-797)."""
    items = list(terms.items())
    n = len(items)
    groups = max(1, min(groups, n)) if n else 0
    out = []
    for g in range(groups):
        lo = n * g // groups
        hi = n * (g + 1) // groups
        if hi > lo:
            out.append(dict(items[lo:hi]))
    return out


def _matvec_job(args):
    terms, n_qubits, x = args
    return matvec(terms, n_qubits, x)


def matvec_parallel(terms, n_qubits, x, processes):
    """ParallelLinearQubitOperator._matvec (linear_qubit_operator.py:183-209):
    term slices in worker processes, full vector to each, numpy.add reduce."""
    pieces = split_terms(terms, processes)
    if not pieces:
        return numpy.zeros(numpy.asarray(x).shape)
    if processes <= 1 or len(pieces) == 1:
        return functools.reduce(numpy.add, (matvec(p, n_qubits, x) for p in pieces))
    ctx = multiprocessing.get_context('fork')
    with ctx.Pool(len(pieces)) as pool:
        parts = pool.imap_unordered(_matvec_job, [(p, n_qubits, x) for p in pieces])
        result = functools.reduce(numpy.add, parts)
    return result


# ----------------------------------------------------------------------------
# a5: get_linear_qubit_operator_diagonal (sparse_tools.py:197-247)
# ----------------------------------------------------------------------------
def diagonal(terms, n_qubits):
    """float64 diagonal; strings containing X or Y contribute nothing.

    The reference builds each Z-string's +-1 vector by recursive split/negate;
    the closed form is (-1)^popc(j & z).  Accumulation is in term order into a
    float64 array, so complex-typed coefficients raise exactly like the
    reference (numpy refuses the float64 += complex cast).
    """
    dim = 1 << n_qubits
    diag = numpy.zeros(dim)
    idx = numpy.arange(dim, dtype=numpy.uint64)
    for term, coefficient in terms.items():
        if any(action in ('X', 'Y') for _, action in term):
            continue
        _, z_mask, _ = term_masks(term, n_qubits)
        signs = 1.0 - 2.0 * bit_parity(idx & numpy.uint64(z_mask)).astype(numpy.float64)
        diag += coefficient * signs
    return diag


# ----------------------------------------------------------------------------
# a6: qubit_operator_sparse (sparse_tools.py:136-194)
# ----------------------------------------------------------------------------
def pauli_term_triplets(term, coefficient, n_qubits):
    """(data, row, col) of coefficient * kron(Paulis) in the order the reference emits.

    sparse_tools.py:180-184: data comes from csc->coo (column-major, entry k is
    column k); csc.nonzero() is row-sorted, and because a Pauli string is an
    involutive signed permutation the two orders pair up as
    (row = k ^ x, col = k, data = M[k ^ x, k]).
    M[c ^ x, c] = coeff * i^y * (-1)^popc(c & z).
    """
    dim = 1 << n_qubits
    x_mask, z_mask, y_count = term_masks(term, n_qubits)
    col = numpy.arange(dim, dtype=numpy.uint64)
    # The values are exact (unit factors +-1, +-i), but the SIGNS OF THEIR ZERO PARTS are not a
    # function of the product alone: scipy.sparse.kron (COO path, `A.data.repeat(B.nnz) * B.data`)
    # multiplies the running value by one complex factor per list entry -- the coefficient, an
    # identity block for every gap, each Pauli matrix (sparse_tools.py:160-178) -- and IEEE
    # complex multiplication turns (-1.5 + 0j) * (1 + 0j) into (-1.5 - 0j).  Replaying the chain
    # reproduces the reference's data bit for bit, signed zeros included.
    factors = {'X': numpy.array([1.0, 1.0], dtype=complex),   # CSC order: column 0, column 1
               'Y': numpy.array([1.0j, -1.0j], dtype=complex),
               'Z': numpy.array([1.0, -1.0], dtype=complex)}
    data = numpy.array([coefficient])
    if data.dtype == object:
        data = data.astype(complex)
    tensor_factor = 0
    for qubit, action in term:
        if qubit > tensor_factor:
            data = (data.reshape(-1, 1) * numpy.ones(1 << (qubit - tensor_factor), dtype=complex)).ravel()
        data = (data.reshape(-1, 1) * factors[action]).ravel()
        tensor_factor = qubit + 1
    if tensor_factor < n_qubits or not term:
        data = (data.reshape(-1, 1) * numpy.ones(1 << (n_qubits - tensor_factor), dtype=complex)).ravel()
    if coefficient == 0:  # coo_matrix(0) has no stored entry: the Kronecker product is empty
        empty = numpy.zeros(0)
        return numpy.zeros(0, dtype=complex), empty, empty
    row = col ^ numpy.uint64(x_mask)
    return data, row.astype(numpy.float64), col.astype(numpy.float64)


def _coo_to_csc(values, rows, cols, dim):
    """sparse_tools.py:187-193 -- concatenation seeded with an empty float list,
    coo -> csc (scipy sums duplicates after a per-column std::sort), eliminate_zeros."""
    values = numpy.concatenate([[]] + values)
    rows = numpy.concatenate([[]] + rows)
    cols = numpy.concatenate([[]] + cols)
    mat = scipy.sparse.coo_matrix((values, (rows, cols)), shape=(dim, dim)).tocsc(copy=False)
    mat.eliminate_zeros()
    return mat


def qubit_operator_sparse(terms, n_qubits=None):
    needed = count_qubits_terms(terms)
    if n_qubits is None:
        n_qubits = needed
    if n_qubits < needed:
        raise ValueError('Invalid number of qubits specified.')
    dim = 1 << n_qubits
    values, rows, cols = [], [], []
    for term, coefficient in terms.items():
        d, r, c = pauli_term_triplets(term, coefficient, n_qubits)
        values.append(d)
        rows.append(r)
        cols.append(c)
    return _coo_to_csc(values, rows, cols, dim)


def qubit_operator_sparse_columns(terms, n_qubits, columns):
    """Columns `columns` (ascending) of qubit_operator_sparse(terms, n_qubits) as a
    (2^n x len(columns)) CSC block, for sizes at which the whole matrix cannot be built on the
    host (BASELINE.json config 2: 7.4e8 non-zeros).  Same pipeline restricted to the block --
    triplets term by term (sparse_tools.py:160-186), scipy's coo -> csc (per-column sort, sum of
    duplicates), eliminate_zeros -- so structure and values are the reference's whenever the
    sums are exact (dyadic coefficients); zero signs use the closed form, not the kron chain."""
    columns = numpy.asarray(columns, dtype=numpy.uint64)
    count = len(terms)
    xm = numpy.zeros(count, dtype=numpy.uint64)
    zm = numpy.zeros(count, dtype=numpy.uint64)
    ph = numpy.zeros(count, dtype=complex)
    for k, (term, coefficient) in enumerate(terms.items()):
        x_mask, z_mask, y_count = term_masks(term, n_qubits)
        xm[k], zm[k] = x_mask, z_mask
        ph[k] = coefficient * _PHASES[y_count % 4]
    keep = ph != 0
    xm, zm, ph = xm[keep], zm[keep], ph[keep]
    rows = columns[None, :] ^ xm[:, None]
    signs = 1.0 - 2.0 * bit_parity(columns[None, :] & zm[:, None]).astype(numpy.float64)
    values = ph[:, None] * signs
    local = numpy.broadcast_to(numpy.arange(len(columns), dtype=numpy.int64)[None, :], rows.shape)
    block = scipy.sparse.coo_matrix((values.ravel(), (rows.ravel().astype(numpy.int64), local.ravel())),
                                    shape=(1 << n_qubits, len(columns))).tocsc(copy=False)
    block.eliminate_zeros()
    return block


# ----------------------------------------------------------------------------
# a7: jordan_wigner_sparse (sparse_tools.py:49-133)
# ----------------------------------------------------------------------------
def ladder_csc(n_qubits, mode, raising):
    """JW ladder operator as CSC: Z_0..Z_{mode-1} (X -+ iY)/2 (sparse_tools.py:49-73).

    a_mode lowers bit b = n-1-mode: columns with the bit set map to the row
    with it cleared; a^dagger is the transpose pattern.  Sign = parity of the
    occupied modes before `mode` (index bits above b).
    """
    dim = 1 << n_qubits
    b = n_qubits - 1 - mode
    idx = numpy.arange(dim, dtype=numpy.uint64)
    bit = numpy.uint64(1 << b)
    above = numpy.uint64(((1 << n_qubits) - 1) ^ ((1 << (b + 1)) - 1))
    sign = 1.0 - 2.0 * bit_parity(idx & above).astype(numpy.float64)
    occupied = (idx & bit) != 0
    cols = idx[~occupied] if raising else idx[occupied]
    rows = cols ^ bit
    data = sign[cols].astype(complex)
    return scipy.sparse.csc_matrix(
        (data, (rows.astype(numpy.int64), cols.astype(numpy.int64))), shape=(dim, dim)
    )


def jordan_wigner_sparse(fermion_terms, n_qubits=None):
    if n_qubits is None:
        n_qubits = count_qubits_terms(fermion_terms)
    dim = 1 << n_qubits
    ladders = [(ladder_csc(n_qubits, q, 0), ladder_csc(n_qubits, q, 1)) for q in range(n_qubits)]
    values, rows, cols = [], [], []
    for term, coefficient in fermion_terms.items():
        mat = coefficient * scipy.sparse.identity(dim, dtype=complex, format='csc')
        for mode, kind in term:
            mat = mat * ladders[mode][kind]
        if coefficient:
            mat = mat.tocoo(copy=False)
            values.append(mat.data)
            r, c = mat.nonzero()
            rows.append(r)
            cols.append(c)
    return _coo_to_csc(values, rows, cols, dim)


# ----------------------------------------------------------------------------
# a9/a10/a15: expectation, variance, inner_product (sparse_tools.py:629-690, 1098)
# ----------------------------------------------------------------------------
def expectation_state(apply_fn, state):
    """<state|H|state> without normalisation; apply_fn(v) -> H v."""
    state = numpy.asarray(state)
    if state.ndim == 1:
        return numpy.dot(numpy.conjugate(state), apply_fn(state))
    return numpy.dot(numpy.conjugate(state.T), apply_fn(state))[0, 0]


def expectation_terms(terms, n_qubits, state):
    return expectation_state(lambda v: matvec(terms, n_qubits, v), state)


def variance_terms(terms, n_qubits, state):
    """<H^2> - <H>^2 with H^2 applied as two matvecs (LinearOperator power)."""
    def apply2(v):
        return matvec(terms, n_qubits, matvec(terms, n_qubits, v))
    return expectation_state(apply2, state) - expectation_terms(terms, n_qubits, state) ** 2


# ----------------------------------------------------------------------------
# a11/a12: get_ground_state / get_gap (sparse_tools.py:566-587, 1078-1095)
# ----------------------------------------------------------------------------
class TermsLinearOperator(scipy.sparse.linalg.LinearOperator):
    def __init__(self, terms, n_qubits):
        dim = 1 << n_qubits
        super().__init__(shape=(dim, dim), dtype=complex)
        self.terms = terms
        self.n_qubits = n_qubits

    def _matvec(self, x):
        return matvec(self.terms, self.n_qubits, x)


def get_ground_state(operator, initial_guess=None):
    values, vectors = scipy.sparse.linalg.eigsh(
        operator, k=1, v0=initial_guess, which='SA', maxiter=1e7
    )
    order = numpy.argsort(values)
    return values[order][0], vectors[:, order][:, 0]


def get_gap(operator, initial_guess=None):
    values, _ = scipy.sparse.linalg.eigsh(operator, k=2, v0=initial_guess, which='SA', maxiter=1e7)
    return abs(values[1] - values[0])


# ----------------------------------------------------------------------------
# a13/a14: Davidson (davidson.py:99-339, 372-469)
# ----------------------------------------------------------------------------
def orthonormalize(vectors, num_orthonormals=1, eps=1e-6):
    """In-place modified Gram-Schmidt with eps-drop (davidson.py:439-469)."""
    kept = num_orthonormals
    for i in range(num_orthonormals, vectors.shape[1]):
        v = vectors[:, i]
        for j in range(i):
            v -= vectors[:, j] * numpy.dot(vectors[:, j].conj(), v)
        if numpy.max(numpy.abs(v)) < eps:
            continue
        vectors[:, kept] = v / numpy.linalg.norm(v)
        kept += 1
    return vectors[:, :kept]


def generate_random_vectors(row, col, real_only=False):
    vecs = numpy.random.rand(row, col)
    if not real_only:
        vecs = vecs + numpy.random.rand(row, col) * 1.0j
    return scipy.linalg.orth(vecs)


def append_random_vectors(vectors, col, max_trial=3, real_only=False):
    if col <= 0:
        return vectors
    have = vectors.shape[1]
    want = min(have + col, vectors.shape[0] + 1)
    trial = 0
    while have < want:
        trial += 1
        vectors = numpy.hstack(
            [vectors, generate_random_vectors(vectors.shape[0], want - have, real_only)]
        )
        vectors = orthonormalize(vectors, have)
        if vectors.shape[1] == have:
            if trial > max_trial:
                warnings.warn('Unable to generate specified number of random vectors',
                              RuntimeWarning)
                break
        else:
            trial = 1
            have = vectors.shape[1]
    return vectors


def davidson_new_directions(diag, eps, error_v, trial_lambda, trial_v):
    """davidson.py:288-339 with the per-element clamp loop vectorised."""
    directions = []
    max_error = 0
    for i in range(error_v.shape[1]):
        r = error_v[:, i]
        if numpy.max(numpy.abs(r)) < eps:
            continue
        max_error = max(max_error, numpy.linalg.norm(r))
        diff = diag - trial_lambda[i]
        dinv = numpy.where(numpy.abs(diff) > eps, 1.0 / numpy.where(diff == 0, 1.0, diff), 1.0 / eps)
        v = trial_v[:, i]
        directions.append(-r + v * numpy.dot(v.conj(), dinv * r) / numpy.dot(v.conj(), dinv * v))
    return directions, max_error


def davidson_iterate(apply_fn, diag, eps, n_lowest, guess_v, guess_mv=None):
    """davidson.py:223-286."""
    if guess_mv is None:
        guess_mv = apply_fn(guess_v)
    dim = guess_v.shape[1]
    if guess_mv.shape[1] < dim:
        guess_mv = numpy.hstack([guess_mv, apply_fn(guess_v[:, guess_mv.shape[1]:dim])])
    vmv = numpy.dot(guess_v.conj().T, guess_mv)
    lam, trans = numpy.linalg.eigh(vmv)
    order = lam.argsort()
    lam, trans = lam[order][:n_lowest], trans[:, order][:, :n_lowest]
    trial_v = numpy.dot(guess_v, trans)
    trial_mv = numpy.dot(guess_mv, trans)
    err = trial_mv - trial_v * lam
    new_dirs, max_err = davidson_new_directions(diag, eps, err, lam, trial_v)
    if new_dirs:
        guess_v = numpy.hstack([guess_v, numpy.stack(new_dirs).T])
    return lam, trial_v, trial_mv, max_err, guess_v, guess_mv


def davidson_lowest_n(apply_fn, diag, n_lowest=1, initial_guess=None, max_iterations=300,
                      max_subspace=100, eps=1e-6, real_only=False):
    """Davidson.get_lowest_n (davidson.py:99-221). apply_fn maps (N,k) -> (N,k)."""
    diag = numpy.asarray(diag)
    max_subspace = min(max_subspace, len(diag) + 1)
    if n_lowest <= 0 or n_lowest >= max_subspace:
        raise ValueError('n_lowest out of range')
    if initial_guess is None:
        initial_guess = generate_random_vectors(len(diag), n_lowest, real_only=real_only)
    if real_only and not numpy.allclose(numpy.real(initial_guess), initial_guess):
        initial_guess = numpy.real(initial_guess)
    if numpy.max(numpy.abs(initial_guess)) < eps:
        raise ValueError('Guess vectors are all zero!')
    guess_v = scipy.linalg.orth(initial_guess)
    if guess_v.shape[1] < n_lowest:
        guess_v = append_random_vectors(guess_v, n_lowest - guess_v.shape[1], real_only=real_only)
    success = False
    guess_mv = None
    it = 0
    while it < max_iterations and not success:
        lam, vecs, mvecs, max_err, guess_v, guess_mv = davidson_iterate(
            apply_fn, diag, eps, n_lowest, guess_v, guess_mv)
        if max_err < eps:
            success = True
            break
        if real_only:
            guess_v = numpy.real(guess_v)
        count_mvs = guess_mv.shape[1]
        guess_v = orthonormalize(guess_v, count_mvs, eps)
        if guess_v.shape[1] <= count_mvs:
            guess_v = append_random_vectors(guess_v, n_lowest, real_only=real_only)
        if guess_v.shape[1] >= max_subspace:
            guess_v = numpy.hstack([vecs, guess_v[:, count_mvs:]])
            guess_mv = mvecs
            if real_only and (not numpy.allclose(numpy.real(guess_v), guess_v)
                              or not numpy.allclose(numpy.real(guess_mv), guess_mv)):
                guess_mv = None
        it += 1
    return success, lam, vecs
