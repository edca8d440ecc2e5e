/* ofb200.h -- C ABI of the B200-native Pauli-sum Hamiltonian engine (libofb200.so).
 *
 * OpenFermion has no FFI; its boundary for this path is the Python API in
 * openfermion/linalg/__init__.py:12-85.  Every entry point below is what a
 * ctypes binding of that API calls; the comment above each one names the
 * reference function (file:line under /root/reference/src/openfermion) whose
 * arithmetic it replaces.  Conventions:
 *   - plain pointers and sizes only, no C++/torch types;
 *   - every function returns an int status (OFB_OK == 0); the message for the
 *     last failure on the calling thread is ofb_last_error();
 *   - "d_" pointers are device pointers on the plan's device, "h_" pointers are
 *     host pointers owned by the caller; nothing allocated here is handed to
 *     the caller except through ofb_malloc/ofb_host_alloc;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - amplitudes are complex128 stored (re, im) = 16 bytes; index bit (n-1-q)
 *     is qubit q (linalg/linear_qubit_operator.py:78-84,126).
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with OFB_ERR_CUDA.
 */
#ifndef OFB200_H
#define OFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_OK 0
#define OFB_ERR_INVALID 1 /* bad argument */
#define OFB_ERR_CUDA 2    /* CUDA runtime error (message has the cudaError string) */
#define OFB_ERR_NOMEM 3   /* host or device allocation failed */
#define OFB_ERR_STATE 4   /* call sequence error (e.g. csc_fill before csc_count) */

typedef struct ofb_plan ofb_plan; /* compiled term tables on one device */

/* ---- library / device ---------------------------------------------------- */
int ofb_version(void);
const char *ofb_last_error(void);
int ofb_device_count(int *count);
int ofb_set_device(int device);
int ofb_device_info(int device, int *sm_count, size_t *total_bytes, size_t *free_bytes,
                    int *cc_major, int *cc_minor);

/* ---- memory / stream / event plumbing (so a ctypes host needs no torch) --- */
int ofb_malloc(void **d_ptr, size_t bytes);
int ofb_free(void *d_ptr);
int ofb_host_alloc(void **h_ptr, size_t bytes); /* pinned */
int ofb_host_free(void *h_ptr);
int ofb_host_register(void *h_ptr, size_t bytes); /* pin caller memory in place */
int ofb_host_unregister(void *h_ptr);
int ofb_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
int ofb_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
int ofb_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream);
/* Blocking copies between device memory and ORDINARY (pageable) host memory through the
 * library's pinned staging buffers and `threads` worker threads (<= 0: $OFB_COPY_THREADS or the
 * cores, at most 16), each on its own stream: the CPU side of the staging -- including the
 * first touch of freshly allocated destination pages -- runs in parallel and overlaps the DMA.
 * device < 0: the calling thread's current device.  Work queued earlier on other streams is NOT
 * waited for: synchronise the producing stream first.
 * What numpy.ndarray results of the reference's calls cost to hand over at 2^28 amplitudes or
 * 7e8 non-zeros (linalg/linear_qubit_operator.py:74-148, linalg/sparse_tools.py:136-194). */
int ofb_copy_h2d_staged(void *d_dst, const void *h_src, size_t bytes, int device, int threads);
int ofb_copy_d2h_staged(void *h_dst, const void *d_src, size_t bytes, int device, int threads);
int ofb_memset(void *d_ptr, int byte, size_t bytes, void *stream);
int ofb_stream_create(void **stream);
int ofb_stream_destroy(void *stream);
int ofb_stream_sync(void *stream);
int ofb_device_sync(void);
int ofb_event_create(void **event);
int ofb_event_destroy(void *event);
int ofb_event_record(void *event, void *stream);
int ofb_event_sync(void *event);
int ofb_event_elapsed_ms(void *start, void *stop, float *ms);
int ofb_stream_wait_event(void *stream, void *event);
/* CUDA IPC, for peer shards of a state vector split over processes */
int ofb_ipc_get_handle(void *d_ptr, void *handle64 /* 64 bytes out */);
int ofb_ipc_open_handle(const void *handle64, void **d_ptr);
int ofb_ipc_close_handle(void *d_ptr);
int ofb_enable_peer_access(int peer_device);
/* Number of kernels launched by this library in this process (bench's gpu_launches). */
int64_t ofb_launch_count(void);

/* ---- plans ----------------------------------------------------------------
 * ofb_plan_create: compile a Pauli sum.  Term t is coeff[t] * P_t with P_t given
 * by its masks exactly as LinearQubitOperator._matvec derives them
 * (linalg/linear_qubit_operator.py:121-134): x_mask = X|Y positions, z_mask =
 * Y|Z positions (so y_count = popc(x & z)).  coeff_ri holds (re, im) pairs of the
 * reference coefficient (the i^y phase is applied inside).  Term order is kept
 * inside each x-group (it is the summation order).  Replaces the per-call mask
 * folding of linear_qubit_operator.py:107-148 and the per-term kron chain of
 * linalg/sparse_tools.py:158-184. */
int ofb_plan_create(int n_qubits, int64_t n_terms, const uint64_t *x_mask, const uint64_t *z_mask,
                    const double *coeff_ri, int device, ofb_plan **plan);

/* ofb_plan_create_projected: compile "projected" rows for the fermion route
 * (linalg/sparse_tools.py:76-133, jordan_wigner_sparse): row t contributes
 * H[c ^ x, c] += coeff * (-1)^popc(c & z) for every column c with
 * (c & occ_mask) == occ_val.  Only the csc_* calls accept such a plan. */
int ofb_plan_create_projected(int n_qubits, int64_t n_rows, const uint64_t *occ_mask,
                              const uint64_t *occ_val, const uint64_t *x_mask,
                              const uint64_t *z_mask, const double *coeff_ri, int device,
                              ofb_plan **plan);
int ofb_plan_destroy(ofb_plan *plan);
/* n_qubits, T (terms kept), G (distinct x-masks), terms in the x==0 group, device */
int ofb_plan_info(const ofb_plan *plan, int *n_qubits, int64_t *n_terms, int64_t *n_groups,
                  int64_t *n_diag_terms, int *device);
/* x-mask of group g in execution order, with its term count (host tables) */
int ofb_plan_group(const ofb_plan *plan, int64_t g, uint64_t *x_mask, int64_t *n_terms);

/* ---- K1: matrix-free H|psi>  (LinearQubitOperator._matvec,
 * linalg/linear_qubit_operator.py:107-148; k columns = scipy _matmat as driven by
 * linalg/davidson.py:249-256).  d_out[:, c] = H d_in[:, c]; columns are `ld`
 * elements apart.  d_in and d_out must not alias. */
int ofb_apply(ofb_plan *plan, const void *d_in, void *d_out, int ncols, int64_t ld, void *stream);

/* Sharded / partial form used by the multi-GPU path and by fused Krylov steps:
 *   out[j] (+)= sum over groups g in [g_begin, g_end) of
 *             S_g(index_base | j) * in[j ^ (x_g & (2^local_bits - 1))]
 * for j in [0, 2^local_bits).  The caller routes the high bits of x_g (which
 * shard `d_in` is).  accumulate != 0 adds into d_out. */
int ofb_apply_groups(ofb_plan *plan, const void *d_in, void *d_out, int local_bits,
                     uint64_t index_base, int64_t g_begin, int64_t g_end, int accumulate,
                     void *stream);

/* Optional "support" hints for K1/K2 (experimental, off unless set): for group g the caller
 * asserts  S_g(j) == 0  unless  parity(j & mask1[g]) == (flags[g] >> 1 & 1)  [if flags[g] & 1]
 *                       and     parity(j & mask2[g]) == (flags[g] >> 3 & 1)  [if flags[g] & 4],
 * j being the global basis index.  Number- and spin-conserving Hamiltonians have such masks for
 * every x != 0 group (a hop X_a X_b + Y_a Y_b only connects states with bit a != bit b); the
 * kernel then skips the gathers of amplitudes whose coefficient is structurally zero (half of
 * them for a hopping group, three quarters for a mixed-spin double excitation).  The hints are
 * trusted: wrong hints give wrong results.  openfermion_b200/support.py derives them from the
 * term table.  n_qubits <= 32 only.  n_groups must equal the plan's group count. */
int ofb_plan_set_support(ofb_plan *plan, int64_t n_groups, const uint32_t *mask1,
                         const uint32_t *mask2, const uint8_t *flags);

/* Host-buffer convenience: H2D, ofb_apply, D2H (the call LinearQubitOperator.matvec makes).
 * h_in/h_out are column-contiguous complex128 arrays of 2^n * ncols elements. */
int ofb_apply_host(ofb_plan *plan, const void *h_in, void *h_out, int ncols);

/* ---- K2: <psi|H|psi> without materialising H psi (expectation,
 * linalg/sparse_tools.py:629-671: numpy.dot(conj(state), operator * state)). */
int ofb_expectation(ofb_plan *plan, const void *d_state, double *h_out_ri /* 2 */, void *stream);
int ofb_expectation_host(ofb_plan *plan, const void *h_state, double *h_out_ri);

/* ---- K3: diagonal (get_linear_qubit_operator_diagonal,
 * linalg/sparse_tools.py:197-247): float64[2^n], only x == 0 terms, real parts. */
int ofb_diagonal(ofb_plan *plan, double *d_out, void *stream);
int ofb_diagonal_host(ofb_plan *plan, double *h_out);

/* ---- K4: CSC materialisation (qubit_operator_sparse linalg/sparse_tools.py:136-194,
 * jordan_wigner_sparse :76-133; replaces scipy coo->csc + sum_duplicates +
 * eliminate_zeros).  Two passes: count (also scans, result = nnz and the index
 * width scipy would choose: 32 iff max(nnz, 2^n) < 2^31), then fill into caller
 * buffers: indptr[2^n + 1], indices[nnz] (int32 or int64 per index_bits) and
 * data[nnz] complex128; rows ascending inside each column, exact zeros dropped. */
int ofb_csc_count(ofb_plan *plan, int64_t *nnz, int *index_bits, void *stream);
/* Summation order of duplicate entries (same row, same column) in the CSC passes:
 *   mode 0  caller's term order (fast; identical to the reference whenever a column holds
 *           <= 16 triplets or the sums are exact);
 *   mode 1  the order scipy produces: coo->csc sorts every column's (row, value) pairs with
 *           libstdc++ std::sort (unstable above 16 elements) before csr_sum_duplicates adds
 *           equal rows left to right (linalg/sparse_tools.py:190-193 via scipy
 *           sparsetools csr_sort_indices / csr_sum_duplicates).  The element movements of
 *           std::sort are replayed per column, so even the floating-point cancellation
 *           residues that survive eliminate_zeros match the reference bit for bit.
 *           Cost O(T log T) per column.  Must be set before ofb_csc_count. */
int ofb_plan_set_csc_mode(ofb_plan *plan, int mode);
/* Signed zeros.  The reference's entries carry the zero signs that scipy.sparse.kron's chain of
 * complex multiplications leaves behind (sparse_tools.py:160-178: coefficient, identity blocks,
 * Pauli matrices, one factor per qubit), e.g. (-1.5 + 0i)(1 + 0i) = -1.5 - 0i.  The CSC kernels
 * replay that chain on a 3-bit state so that `data` matches bit for bit; this host function is
 * the same replay for one entry (string x_mask / z_mask with coefficient (re, im), column
 * `column`): *negative = 1 if the zero part of the entry is -0.0, *has_zero = 0 if neither or
 * both parts are zero (nothing to sign). */
int ofb_pauli_entry_zero_sign(int n_qubits, uint64_t x_mask, uint64_t z_mask, double re, double im,
                              uint64_t column, int *has_zero, int *negative);
int ofb_csc_fill(ofb_plan *plan, void *d_indptr, void *d_indices, void *d_data, int index_bits,
                 void *stream);
int ofb_csc_fill_host(ofb_plan *plan, void *h_indptr, void *h_indices, void *h_data,
                      int index_bits);

/* CSR/CSC-stored operator times vector for get_ground_state / SparseDavidson on an
 * already materialised matrix (linalg/sparse_tools.py:578, linalg/davidson.py:369):
 * y = A x with A in CSR (indptr[n_rows+1], indices, data complex128). */
int ofb_csr_matvec(int64_t n_rows, const void *d_indptr, const void *d_indices, const void *d_data,
                   int index_bits, const void *d_x, void *d_y, void *stream);

/* ---- K5: vector kernels for the device-resident Krylov loops
 * (numpy calls in linalg/davidson.py:257-339,439-469 and ARPACK internals behind
 * linalg/sparse_tools.py:578).  All vectors complex128 of length n. Scalars that
 * come back are written to host memory after a stream synchronise.
 * Threading: the reductions (dotc, nrm2, maxabs, multi_dotc, davidson_correction) share ONE
 * scratch buffer and one pinned result buffer per device; each call holds a per-device lock from
 * its first kernel to the stream synchronisation that ends it, so calls from several host threads
 * or streams serialise.  The plan / sector reductions (expectation) use per-object scratch: one
 * thread and stream at a time per plan (as a numpy caller of the reference: its matvec is
 * single-threaded). */
int ofb_vec_dotc(int64_t n, const void *d_x, const void *d_y, double *h_out_ri, void *stream);
int ofb_vec_nrm2(int64_t n, const void *d_x, double *h_out, void *stream);
int ofb_vec_maxabs(int64_t n, const void *d_x, double *h_out, void *stream);
int ofb_vec_axpy(int64_t n, const double *alpha_ri, const void *d_x, void *d_y, void *stream);
int ofb_vec_scal(int64_t n, const double *alpha_ri, void *d_x, void *stream);
/* out[k] = conj(V[:,k]) . w for k < ncols (columns ld apart): V^H w in one pass over w */
int ofb_vec_multi_dotc(int64_t n, int ncols, const void *d_V, int64_t ld, const void *d_w,
                       double *h_out_ri /* 2*ncols */, void *stream);
/* w = beta*w + sum_k a[k] V[:,k] */
int ofb_vec_multi_axpy(int64_t n, int ncols, const double *a_ri /* 2*ncols */, const void *d_V,
                       int64_t ld, const double *beta_ri, void *d_w, void *stream);
/* Lanczos step tail: w -= alpha*v + beta*v_prev (alpha, beta real) */
int ofb_vec_lanczos_update(int64_t n, double alpha, double beta, const void *d_v,
                           const void *d_vprev, void *d_w, void *stream);
/* Davidson correction (linalg/davidson.py:321-336): with dinv[j] = 1/(diag[j]-lambda) if
 * |diag[j]-lambda| > eps else 1/eps, writes d_new = -r + v * (v^H dinv r)/(v^H dinv v). */
int ofb_davidson_correction(int64_t n, const double *d_diag, double lambda, double eps,
                            const void *d_r, const void *d_v, void *d_new, void *stream);
/* real parts only (DavidsonOptions.real_only, linalg/davidson.py:183-184) */
int ofb_vec_drop_imag(int64_t n, void *d_x, void *stream);
/* deterministic pseudo-random fill in [0,1) (+ i[0,1) unless real_only), counter based */
int ofb_vec_fill_random(int64_t n, uint64_t seed, int real_only, void *d_x, void *stream);

/* ---- K7: symmetry sectors (particle number / S_z) --------------------------
 * The reference restricts an operator by materialising the full 2^n x 2^n matrix and slicing it
 * with numpy.ix_ (linalg/sparse_tools.py:375-423 jw_number_restrict_operator /
 * jw_sz_restrict_operator on the index lists of :250-371 jw_number_indices / jw_sz_indices), and
 * builds a sector matrix term by term in get_number_preserving_sparse_operator (:1282-1597).
 * Here a sector is an index space of `dim` basis states; position p holds basis state basis[p]
 * (reference bit convention) in the reference's own enumeration order, and the kernels below act
 * on vectors of `dim` complex128 amplitudes. */
typedef struct ofb_sector ofb_sector;

/* Two-table index map: for j = (hi, lo) split at lo_bits,
 *   pos(j) = table_hi[hi * n_classes + popc(lo & class_mask)]
 *          + table_lo[lo * n_classes + popc(hi-part of j & class_mask)]
 * (n_classes == 1: the class index is 0).  j belongs to the sector iff for every constraint k
 * popc(j & plus_mask[k]) - popc(j & minus_mask[k]) == value[k].  The basis list is enumerated on
 * the device.  Replaces the itertools enumeration of sparse_tools.py:250-371. */
int ofb_sector_create_tables(int n_qubits, int lo_bits, int n_classes, uint64_t class_mask,
                             const uint64_t *table_hi, const uint32_t *table_lo, int n_constraints,
                             const uint64_t *plus_mask, const uint64_t *minus_mask,
                             const int32_t *value, int64_t dim, int device, ofb_sector **sector);
/* Arbitrary basis list in the caller's order (custom spin maps, the excitation-ordered bases of
 * sparse_tools.py:1370-1475 _iterate_basis_): lookups are a binary search (the reference uses
 * numpy.searchsorted, :1573-1584). */
int ofb_sector_create_list(int n_qubits, int64_t dim, const uint64_t *h_basis, int device,
                           ofb_sector **sector);
int ofb_sector_destroy(ofb_sector *sector);
int ofb_sector_info(const ofb_sector *sector, int *n_qubits, int64_t *dim, int *mode);
/* basis state at every position (jw_number_indices / jw_sz_indices as an array) */
int ofb_sector_basis(ofb_sector *sector, uint64_t *h_basis /* dim */);
/* Matrix-free (P H P)|psi> inside the sector: the action of
 * jw_number_restrict_operator(get_sparse_operator(H)) / jw_sz_restrict_operator(...) on a vector
 * without building either matrix.  ncols columns `ld` elements apart. */
int ofb_sector_apply(ofb_plan *plan, ofb_sector *sector, const void *d_in, void *d_out, int ncols,
                     int64_t ld, void *stream);
int ofb_sector_apply_host(ofb_plan *plan, ofb_sector *sector, const void *h_in, void *h_out,
                          int ncols);
int ofb_sector_expectation(ofb_plan *plan, ofb_sector *sector, const void *d_state,
                           double *h_out_ri /* 2 */, void *stream);
/* diagonal of the restricted operator, float64[dim] (Davidson preconditioner) */
int ofb_sector_diagonal(ofb_plan *plan, ofb_sector *sector, double *d_out, void *stream);
/* jw_number_restrict_state / jw_sz_restrict_state (sparse_tools.py:426-476): sector[p] =
 * full[basis[p]]; and the expansion of :507-509: full = 0, full[basis[p]] = sector[p]. */
int ofb_sector_gather(ofb_sector *sector, const void *d_full, void *d_sector, void *stream);
int ofb_sector_scatter(ofb_sector *sector, const void *d_sector, void *d_full, void *stream);
/* CSC of the restricted operator (dim x dim), same contract as ofb_csc_count / ofb_csc_fill_host;
 * accepts Pauli plans and projected (fermion-route) plans. */
int ofb_sector_csc_count(ofb_plan *plan, ofb_sector *sector, int64_t *nnz, int *index_bits,
                         void *stream);
int ofb_sector_csc_fill_host(ofb_plan *plan, ofb_sector *sector, void *h_indptr, void *h_indices,
                             void *h_data, int index_bits);

/* ---- N3: fermion-to-qubit term generation on the device (csrc/termgen.cu) ----------------------
 * Replaces the symbolic products of transforms/opconversions/jordan_wigner.py:132-331 and
 * bravyi_kitaev.py:242-337 for operators given as ladder-operator products.  An encoding is four
 * per-mode mask tables: a_j^(dagger) = 1/2 (A_j -/+ i B_j) with A_j = P(xa[j], za[j]),
 * B_j = P(xb[j], zb[j]) (qubit-space masks, bit q = qubit q; Jordan-Wigner and Bravyi-Kitaev
 * tables come from openfermion_b200/hamiltonians.py).  ofb_termgen_add expands
 * sum_t coeff[t] prod_j ladder(modes[t][j], kind j) -- k factors per term, bit j of kinds_mask = 1
 * for a raising operator -- into its 2^k strings per term; ofb_termgen_collect merges equal
 * strings (stable sort + sequential sums: the host builder's summation order, bit-identical
 * coefficients), drops |sum| < tol (the reference's 1e-8 rule, symbolic_operator.py __iadd__) and
 * leaves the strings sorted by (x, z); ofb_termgen_fetch copies them out. */
typedef struct ofb_termgen ofb_termgen;
int ofb_termgen_create(int n_qubits, const uint64_t *xa, const uint64_t *za, const uint64_t *xb,
                       const uint64_t *zb, int device, ofb_termgen **gen);
int ofb_termgen_add(ofb_termgen *gen, int k, uint32_t kinds_mask, int64_t n_terms,
                    const int32_t *h_modes /* n_terms * k */, const double *h_coeff_ri /* 2 n_terms */);
int ofb_termgen_collect(ofb_termgen *gen, double tol, int64_t *n_out);
int ofb_termgen_fetch(ofb_termgen *gen, uint64_t *h_x, uint64_t *h_z, double *h_coeff_ri);
int ofb_termgen_destroy(ofb_termgen *gen);

/* ---- GF(2) relabelling of basis indices (csrc/relabel.cu, openfermion_b200/shard_basis.py) -------
 * h_cols[b] (b < n_bits <= 40) is the image of index bit b; the image of an index is the XOR of
 * the images of its set bits.  The reference keeps qubit q at index bit n-1-q
 * (linalg/linear_qubit_operator.py:78-84); relabelled solves convert on the way in and out.
 * ofb_vec_fill_random_mapped: element i gets the counter-based random value of index
 * map(index_base | i) (the generator of ofb_vec_fill_random): a shard of ONE global random state
 * whatever the sharding or basis.  ofb_vec_relabel: d_out[c] = d_in[map(c)] over 2^n_bits. */
int ofb_vec_fill_random_mapped(int64_t n, uint64_t seed, int real_only, void *d_x,
                               uint64_t index_base, int n_bits, const uint64_t *h_cols, void *stream);
int ofb_vec_relabel(int n_bits, const uint64_t *h_cols, const void *d_in, void *d_out, void *stream);

/* ---- device-resident Lanczos (csrc/krylov.cu) --------------------------------------------------
 * Replaces the ARPACK call of get_ground_state / get_gap (linalg/sparse_tools.py:578,1090:
 * eigsh(k=1|2, which='SA')) with a restarted three-term Lanczos whose vectors never leave the
 * device.  An ofb_linop is a handle on "y = A x" for one device: a compiled plan (K1), a plan
 * inside a symmetry sector (K7) or a CSR matrix already on the device (ofb_csr_matvec). */
typedef struct ofb_linop ofb_linop;
int ofb_linop_from_plan(ofb_plan *plan, ofb_linop **op);
int ofb_linop_from_sector(ofb_plan *plan, ofb_sector *sector, ofb_linop **op);
int ofb_linop_from_csr(int64_t n_rows, const void *d_indptr, const void *d_indices,
                       const void *d_data, int index_bits, int device, ofb_linop **op);
int ofb_linop_destroy(ofb_linop *op);
int ofb_linop_dim(const ofb_linop *op, int64_t *dim);
int ofb_linop_apply(ofb_linop *op, const void *d_in, void *d_out, void *stream);
/* Lowest eigenpair of a Hermitian operator, orthogonal to the n_locked unit vectors d_locked[]
 * (get_gap's second state).  d_start: device start vector (any norm) or NULL for a seeded random
 * one.  Sweeps of at most krylov_dim steps (<= 0: 400) are restarted from the Ritz vector until
 * the TRUE residual |A y - theta y| <= 10 tol max(1, |theta|); *h_energy is the Rayleigh quotient
 * of the returned unit vector d_vector (dim complex128).  store_basis != 0: keep the Lanczos
 * vectors in device memory when they fit (m instead of 2m - 1 applications of A per sweep). */
int ofb_lanczos_lowest(ofb_linop *op, const void *d_start, int n_locked, const void *const *d_locked,
                       double tol, int krylov_dim, int max_restarts, uint64_t seed, int store_basis,
                       double *h_energy, double *h_residual, void *d_vector, int64_t *h_matvecs,
                       int *h_restarts, void *stream);
/* Lowest eigenpair of a real symmetric tridiagonal matrix (bisection + inverse iteration, what
 * scipy.linalg.eigh_tridiagonal(select='i') does through LAPACK): the Ritz problem of a sweep. */
int ofb_tridiag_lowest(int m, const double *diag, const double *offdiag, double *value,
                       double *vector);
/* Per-iteration primitives on UNNORMALISED Lanczos vectors u_k = beta_k v_k, all scalars on the
 * device (no host synchronisation inside an iteration); the sharded solver
 * (openfermion_b200/distributed.py) all-reduces scalar slots 0..1 (the dot product) and 2 (the
 * squared norm) over the ranks between the calls:
 *   ofb_apply_groups_dot / ofb_kws_dot      slots 0,1 = conj(d_dot) . out  (local part)
 *   ofb_kws_alpha(k)                        alpha_k = slot0 / beta_k^2 and the update coefficients
 *   ofb_kws_step                            w <- w/beta_k - (alpha_k/beta_k) u - (beta_k/beta_{k-1}) u_prev
 *                                           [y += (ritz_k/beta_k) u], slot 2 = local |w|^2
 *   ofb_kws_beta(k)                         beta_{k+1} = sqrt(slot 2)
 * replay != 0 repeats a recorded sweep from its stored alpha/beta (Ritz-vector pass). */
typedef struct ofb_kws ofb_kws;
int ofb_kws_create(int64_t n_local, int max_steps, int device, ofb_kws **ws);
int ofb_kws_destroy(ofb_kws *ws);
int ofb_kws_scalars(ofb_kws *ws, void **d_scalars /* 16 doubles on the device */);
int ofb_kws_reset(ofb_kws *ws, double beta0, void *stream);
int ofb_kws_dot(ofb_kws *ws, const void *d_x, const void *d_y, void *stream);
int ofb_apply_groups_dot(ofb_plan *plan, const void *d_in, void *d_out, int local_bits,
                         uint64_t index_base, int64_t g_begin, int64_t g_end, int accumulate,
                         const void *d_dot, ofb_kws *ws, void *stream);
int ofb_kws_alpha(ofb_kws *ws, int k, int replay, double ritz_k, void *stream);
int ofb_kws_step(ofb_kws *ws, void *d_w, const void *d_u, const void *d_uprev, void *d_y, int replay,
                 void *stream);
int ofb_kws_beta(ofb_kws *ws, int k, void *stream);
int ofb_kws_fetch(ofb_kws *ws, int count, double *h_alphas, double *h_betas, int *h_flag,
                  void *stream);

/* ---- Davidson's method (csrc/davidson.cu) --------------------------------------------------------
 * Davidson.get_lowest_n of the reference (linalg/davidson.py:99-221 with _iterate :223-286,
 * _get_new_directions :288-339, orthonormalize :439-469, append_random_vectors :390-436) as ONE
 * native call: guess_v / guess_mv are device column blocks, the operator an ofb_linop, the
 * projected m x m Hermitian problem is solved on the host (ofb_hermitian_eigh), nothing 2^n-long
 * leaves the device.  d_diag: the operator's real diagonal (dim doubles, device) for the
 * clamped-inverse preconditioner; d_guess: n_guess start columns (dim complex128 each, contiguous;
 * n_guess may be 0: seeded random columns); max_subspace / max_iterations / eps / real_only are
 * DavidsonOptions' fields (linalg/davidson.py:34-68; max_subspace is capped at dim + 1 as
 * set_dimension does and at what fits in 45 % of the free device memory).  Outputs:
 * h_eigenvalues[n_lowest] ascending, d_eigenvectors dim x n_lowest column-major (device),
 * *h_success (max residual norm < eps), *h_iterations, *h_max_error (largest residual norm of the
 * last iteration), *h_flags bit 0: fewer random vectors than requested could be generated
 * (the reference's RuntimeWarning). */
int ofb_davidson_lowest_n(ofb_linop *op, const double *d_diag, int n_lowest, int n_guess,
                          const void *d_guess, int max_subspace, int max_iterations, double eps,
                          int real_only, uint64_t seed, double *h_eigenvalues, void *d_eigenvectors,
                          int *h_success, int *h_iterations, double *h_max_error, int *h_flags,
                          void *stream);
/* Eigen-decomposition of a complex Hermitian m x m matrix on the host (numpy.linalg.eigh in
 * linalg/davidson.py:263): a column-major (re, im) pairs, lower triangle read; w[m] ascending,
 * z column-major orthonormal eigenvectors.  Householder tridiagonalisation + implicit QL. */
int ofb_hermitian_eigh(int m, const double *a_ri, double *w, double *z_ri);

#ifdef __cplusplus
}
#endif
#endif /* OFB200_H */
