// Internal declarations shared by the translation units of libofb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "ofb200.h"

namespace ofb {

// ---- error plumbing --------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
extern std::atomic<int64_t> g_launches;

#define OFB_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ofb::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// No C++ exception may cross the C ABI: entry points that allocate host memory wrap their body
// in OFB_TRY / OFB_CATCH(cleanup) and turn std::bad_alloc & co. into a status + message.
#define OFB_TRY try {
#define OFB_CATCH(cleanup)                                                                  \
    }                                                                                       \
    catch (const std::bad_alloc &) {                                                        \
        cleanup;                                                                            \
        return ofb::fail(OFB_ERR_NOMEM, "host allocation failed");                          \
    }                                                                                       \
    catch (const std::exception &e__) {                                                     \
        cleanup;                                                                            \
        return ofb::fail(OFB_ERR_STATE, std::string("internal error: ") + e__.what());      \
    }

#define OFB_LAUNCH_CHECK()                              \
    do {                                                \
        ofb::g_launches.fetch_add(1);                   \
        OFB_CUDA(cudaGetLastError());                   \
    } while (0)

// ---- device tables ---------------------------------------------------------
// One x-group of the gather form  out[j] += S_g(j) * in[j ^ x],
// S_g(j) = sum_t c_t (-1)^popc(j & z_t)  (SURVEY.md section 0.4).
enum GroupKind : uint32_t { KIND_COMPLEX = 0, KIND_REAL = 1, KIND_IMAG = 2 };
struct __align__(16) GroupDesc {  // scatter-form groups of the CSC builder
    uint64_t x;
    uint32_t begin;  // first row
    uint32_t count;  // bits 0..29 row count
};
// Gather-form execution unit: at most 255 terms of one x-group, sorted by the "class"
// w = (z >> CLASS_SHIFT) & CLASS_MASK, i.e. by the z bits that distinguish the 2^CLASS_BITS
// amplitudes a thread owns (at most SEG_MAX_RUNS class runs per segment).  Only the non-empty classes are listed: nibble k of `ids` is the class of the
// k-th run of terms and byte k of `cnts` its length.
#ifndef OFB_APPLY_BLOCK_BITS
#define OFB_APPLY_BLOCK_BITS 6
#endif
#ifndef OFB_APPLY_CLASS_BITS
#define OFB_APPLY_CLASS_BITS 4  // a thread owns 2^CLASS_BITS amplitudes (3 or 4)
#endif
#ifndef OFB_APPLY_MIN_CTAS  // 64-thread CTAs: 6 per SM at 167 registers (16 amplitudes), 8 at 113 (8)
#define OFB_APPLY_MIN_CTAS (OFB_APPLY_CLASS_BITS == 3 ? 8 : 6)
#endif
constexpr int CLASS_SHIFT = OFB_APPLY_BLOCK_BITS;  // log2(threads per block) of the apply kernel
constexpr int CLASS_BITS = OFB_APPLY_CLASS_BITS;
constexpr uint32_t CLASS_MASK = (1u << CLASS_BITS) - 1u;
constexpr int SEG_MAX_RUNS = 8;  // nibbles of Segment::ids / bytes of Segment::cnts
static_assert(CLASS_BITS == 3 || CLASS_BITS == 4, "class ids are stored in nibbles");
constexpr int SEG_MAX_TERMS = 255;
struct __align__(16) Segment {
    uint32_t x_lo;       // low 32 bits of the x mask (all a device-local index can use)
    uint32_t rel_begin;  // first term, relative to the first term of the segment's chunk
    uint32_t kind_runs;  // bits 0..3 number of class runs, bits 8..23 term count, 30..31 kind
    uint32_t ids;        // nibble k = class of run k
    uint64_t cnts;       // byte k = length of run k
    uint64_t x;          // full mask (host side, small-state kernel)
};
// Staging unit of the apply kernel: consecutive segments whose descriptors and term tables
// are copied to shared memory together (double buffered with cp.async).
constexpr int CHUNK_MAX_SEGS = 16;
constexpr int CHUNK_MAX_TERMS = 256;
struct __align__(16) Chunk {
    uint32_t seg_begin, seg_count;
    uint32_t term_begin, term_count;
};
// Optional support hint of a segment's group (ofb_plan_set_support): amplitude r of a thread
// whose r = 0 amplitude has global index j needs the group only if
//   parity(r & w_k) == v_k ^ parity(j & m_k)   for k = 1, 2,   w_k = (m_k >> CLASS_SHIFT) & CLASS_MASK.
// `pat` holds the two 16-bit patterns {r : parity(r & w_k) = 1} (low half k = 1), `v` bit k-1 = v_k.
// A missing condition is stored as m = 0, v = 1, pattern 0xffff (always needed).
struct __align__(16) SegCond {
    uint32_t m1, m2;
    uint32_t pat;
    uint32_t v;
};
// One component (real or imaginary part) of c_t = coeff_t * (-i)^y_t next to the z mask, so
// that one 128-bit warp-uniform load fetches both.
struct __align__(16) TermZC {
    uint64_t z;
    double c;
};
// Scatter-form row for CSC assembly: H[col ^ x, col] += c * (-1)^popc(col & z)
// for columns with (col & occ_mask) == occ_val.  x lives in the group.
struct __align__(16) ScatterRow {
    uint64_t z;
    uint64_t occ_mask;
    uint64_t occ_val;
    uint64_t x;
    double cr, ci;
    uint64_t flags;  // signed-zero replay, see below
};

// ---- signed zeros of the reference's matrix entries ----------------------------------------------
// A matrix entry of qubit_operator_sparse is exact (coefficient times +-1 / +-i), but the SIGN OF
// ITS ZERO PART is not: scipy.sparse.kron multiplies the running value by one complex factor per
// qubit (identity blocks, Pauli matrices; linalg/sparse_tools.py:160-178), and IEEE complex
// multiplication makes (-1.5 + 0i)(1 + 0i) = -1.5 - 0i.  To reproduce the reference's `data`
// bit for bit the kernels replay that chain on a 3-bit state
//     bit 2: which part is zero (0: imaginary, 1: real), bit 1: sign of the non-zero part,
//     bit 0: sign of the zero part,
// with one transition table per factor value (8 states x 3 bits, generated from numpy's own
// complex multiplication by oracle/make_golden.py `signed_zero_tables`): E = 1+0i (identity, X,
// Z on a 0 bit), N = -1+0i (Z on a 1 bit), Y0 = +0+1i / Y1 = -0-1i (Y on a 0 / 1 bit of the column).
// E is idempotent, so one E per gap between non-E factors equals the reference's identity blocks.
// Whenever the chain ends in E and starts from a state a Python float or an ordinary complex
// number produces, the zero part is +0: only strings acting on the LAST qubit need the replay.
constexpr uint32_t SZ_E = 0xdac680u, SZ_N = 0x937052u, SZ_Y0 = 0x212de4u, SZ_Y1 = 0x48193eu;
constexpr uint64_t SZ_STATE = 7, SZ_HAS_ZERO = 8, SZ_NEED_REPLAY = 16;

__host__ __device__ inline uint32_t sz_step(uint32_t table, uint32_t state) {
    return (table >> (3u * state)) & 7u;
}
// state after the chain of column c (qubit 0 = index bit n-1 first)
__host__ __device__ inline uint32_t sz_replay(uint32_t state, uint64_t x, uint64_t z, uint64_t c, int n) {
    uint64_t active = z & (x | c);  // Y anywhere, Z on a set bit: the factors that are not E
    int prev = n;
    while (active) {
#ifdef __CUDA_ARCH__
        const int b = 63 - __clzll((long long)active);
#else
        const int b = 63 - __builtin_clzll(active);
#endif
        if (b < prev - 1) state = sz_step(SZ_E, state);
        const uint32_t table = ((x >> b) & 1ull) ? (((c >> b) & 1ull) ? SZ_Y1 : SZ_Y0) : SZ_N;
        state = sz_step(table, state);
        prev = b;
        active &= ~(1ull << b);
    }
    if (prev > 0) state = sz_step(SZ_E, state);
    return state;
}
// flags of a row from the ORIGINAL coefficient (as numpy sees it) and the string's z mask
__host__ __device__ inline uint64_t sz_flags(double re, double im, uint64_t z, bool pauli_route) {
    const bool zr = re == 0.0, zi = im == 0.0;
    if (zr == zi) return 0;  // both parts non-zero (no zero to sign) or the term is zero
    if (!pauli_route) return SZ_HAS_ZERO;  // fermion route: sparse products always leave +0
    const uint32_t sr = re < 0.0 || (zr && 1.0 / re < 0.0), si = im < 0.0 || (zi && 1.0 / im < 0.0);
    const uint32_t state = zi ? ((sr << 1) | si) : (4u | (si << 1) | sr);
    const bool ordinary = state != 3u && state != 5u;  // (-a, -0) and (-0, +b) break the shortcut
    return SZ_HAS_ZERO | state | ((!ordinary || (z & 1ull)) ? SZ_NEED_REPLAY : 0);
}

#ifdef __CUDACC__
// value of a scatter row in column c: +-(cr, ci) with the reference's signed zeros
__device__ __forceinline__ double2 scatter_value(const ScatterRow &r, uint64_t c, int n) {
    const int sgn = (int)((uint32_t)__popcll(c & r.z) << 31);
    double re = __hiloint2double(__double2hiint(r.cr) ^ sgn, __double2loint(r.cr));
    double im = __hiloint2double(__double2hiint(r.ci) ^ sgn, __double2loint(r.ci));
    if (r.flags & SZ_HAS_ZERO) {
        uint32_t negative = 0;
        if (r.flags & SZ_NEED_REPLAY) negative = sz_replay((uint32_t)(r.flags & SZ_STATE), r.x, r.z, c, n) & 1u;
        const double zero = negative ? -0.0 : 0.0;
        if (re == 0.0) re = zero; else im = zero;
    }
    return make_double2(re, im);
}
#endif

}  // namespace ofb

struct ofb_plan {
    int n_qubits = 0;
    int device = 0;
    bool projected = false;  // fermion-route rows: only CSC calls are valid
    int64_t n_terms = 0;
    int64_t n_groups = 0;
    int64_t n_diag_terms = 0;
    bool z_fits32 = false;
    // host copies (group order = ascending x)
    std::vector<uint64_t> h_group_x;
    std::vector<uint32_t> h_group_begin;  // n_groups + 1
    std::vector<uint32_t> h_group_kind;
    // device tables: gather form (apply / expectation / diagonal)
    std::vector<uint32_t> h_seg_begin;  // n_groups + 1: first segment of each logical group
    std::vector<uint32_t> h_chunk_of_seg;  // chunk holding each segment
    std::vector<ofb::Chunk> h_chunks;      // host copy of d_chunks (slice descriptors of the CSC passes)
    int64_t n_chunks = 0;
    bool has_imag = false;  // some group has a non-real phased coefficient
    ofb::Chunk *d_chunks = nullptr;
    ofb::Segment *d_segs = nullptr;
    uint32_t *d_seg_term_begin = nullptr;  // absolute first term of each segment (small-state kernel)
    ofb::TermZC *d_zc = nullptr;     // (z, Re c) in segment/class order
    ofb::TermZC *d_zc_im = nullptr;  // (z, Im c) in the same order
    ofb::GroupDesc *d_groups = nullptr;  // scatter-form groups (CSC)
    ofb::SegCond *d_conds = nullptr;     // per segment, only after ofb_plan_set_support
    bool has_cond = false;
    // device tables: scatter form (CSC)
    ofb::ScatterRow *d_rows = nullptr;       // grouped by x (stable), for the fast CSC mode
    ofb::ScatterRow *d_rows_orig = nullptr;  // caller's term order, for the scipy-order mode
    int csc_mode = 0;                        // 0 = term-order sums, 1 = scipy sort-order emulation
    uint64_t *d_sort_scratch = nullptr;      // exact mode: [n_terms][batch] (row << 32 | term)
    int64_t sort_batch = 0;                  // columns per batch; == 2^n when one batch suffices
    uint32_t *d_sib = nullptr;  // [n_groups][n_qubits] sibling-subtree sizes
    // tiled CSC passes (csc_tile.cu): group of every segment, groups that need the per-row path
    uint32_t *d_seg_group = nullptr;
    uint8_t *d_group_slow = nullptr;
    std::vector<uint8_t> h_group_slow;
    // slices of the group list for the tiled CSC passes: chunk, segment and term range per slice;
    // the count and fill passes choose their own number
    void *d_count_slices = nullptr, *d_fill_slices = nullptr;  // CscSlice[] (csc_tile.cu)
    uint32_t n_count_slices = 0, n_fill_slices = 0;
    // CSC state between count and fill
    int64_t *d_colptr = nullptr;   // 2^n + 1 exclusive scan of counts
    uint32_t *d_bitmap = nullptr;  // [words][2^n] rank-ordered non-zero flags
    int64_t csc_nnz = -1;
    int64_t bitmap_words = 0;
    // scratch for reductions
    double *d_partials = nullptr;
    size_t partials_bytes = 0;
    double *h_pinned_scalars = nullptr;  // 64 doubles
    // staging for *_host calls
    void *d_stage_in = nullptr;
    void *d_stage_out = nullptr;
    size_t stage_in_bytes = 0, stage_out_bytes = 0;
    cudaStream_t copy_stream = nullptr;  // ofb_apply_host: slab-wise D2H behind the kernel
    cudaEvent_t slab_done[8] = {};
};

namespace ofb {
int ensure_partials(ofb_plan *plan, size_t bytes);
int ensure_stage(ofb_plan *plan, size_t in_bytes, size_t out_bytes);
// launches implemented in apply.cu
int launch_apply(ofb_plan *plan, const void *d_in, void *d_out, int local_bits, uint64_t index_base,
                 int64_t g0, int64_t g1, int accumulate, cudaStream_t s);
int launch_apply_dot(ofb_plan *plan, const void *d_in, void *d_out, int local_bits, uint64_t index_base,
                     int64_t g0, int64_t g1, int accumulate, const void *d_dot, double *d_partials,
                     int64_t *n_partials, cudaStream_t s);
int64_t apply_partial_count(int local_bits);
int launch_expectation(ofb_plan *plan, const void *d_state, double *h_out, cudaStream_t s);
int launch_diagonal(ofb_plan *plan, double *d_out, cudaStream_t s);
// csc_exact.cu: scipy-order (std::sort emulation) variants of the two CSC passes
int csc_exact_count(ofb_plan *plan, int64_t *d_counts, cudaStream_t s);
int csc_exact_fill(ofb_plan *plan, const int64_t *d_colptr, void *d_indices, void *d_data,
                   int index_bits, cudaStream_t s);
// csc_tile.cu: tiled count / fill passes for Pauli plans (term tables shared by 1 024 columns)
// hostcopy.cu: blocking copy between a device pointer and pageable host memory through pinned
// staging buffers and worker threads (threads <= 0: OFB_COPY_THREADS or the cores, at most 16)
int staged_copy(void *d, void *h, size_t bytes, bool to_device, int device, int threads);
bool csc_tile_ok(const ofb_plan *plan);
int csc_tile_count(ofb_plan *plan, uint32_t words, uint32_t *bitmap, uint32_t *prefix, int64_t *counts,
                   cudaStream_t s);
int csc_tile_fill(ofb_plan *plan, uint32_t words, uint32_t *bitmap, uint32_t *prefix,
                  const void *d_indptr /* index_bits wide, already converted */, void *d_indices,
                  void *d_data, int index_bits, cudaStream_t s);
// shared reduction helper (vec.cu): sums `count` rows of `width` doubles
int reduce_partials(const double *d_partials, int64_t count, int width, double *d_out,
                    cudaStream_t s);
double *global_scratch(size_t bytes);  // per-device growable scratch for vec kernels
double *global_pinned();               // 4096 pinned doubles
}  // namespace ofb
