// K7: symmetry-sector (particle number / S_z) restricted forms of K1, K3 and K4.
//
// Reference: linalg/sparse_tools.py:250-371 (jw_number_indices, jw_sz_indices),
// :375-513 (jw_*_restrict_operator/state, jw_get_ground_state_at_particle_number) and
// :1282-1597 (get_number_preserving_sparse_operator).  The reference materialises the full
// 2^n x 2^n matrix and slices it with numpy.ix_; here the sector is a first-class index
// space: position p of the sector holds basis state basis[p] (an n-bit index in the
// reference's bit convention) and every kernel works on vectors of `dim` amplitudes.
//
// Index map (position of a basis state), two forms:
//   table mode   pos(j) = A[hi * K + c_lo] + B[lo * K + c_hi],  j = (hi, lo) split at lo_bits,
//                c_lo / c_hi = popcounts of the class mask in each half (K = 1: fixed-number
//                sectors, the class is ignored).  The host fills A and B so that pos() is the
//                reference's own enumeration order (openfermion_b200/sectors.py); membership
//                is a list of constraints popc(j & plus) - popc(j & minus) == value.
//   search mode  arbitrary basis lists (custom spin maps, excitation-truncated spaces):
//                binary search in the sorted basis, then a permutation to the caller's order.
//
// Matrix-free apply (gather form, the plan's k_apply1 tables):
//   out[p] = sum_g [basis[p] ^ x_g in sector] S_g(basis[p]) * in[pos(basis[p] ^ x_g)]
// one thread per position; the table reads are warp-uniform.  With the reference's
// up-major S_z order a spin-up hop moves whole rows of the (up, down) grid, so the gathers
// stay coalesced.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ofb_internal.cuh"

namespace ofb {

constexpr int SECTOR_MAX_CON = 4;
constexpr int SBLOCK = 256;
#ifndef OFB_SECTOR_BATCH  // segments looked up together; measured 1 / 2 / 4: 25.7 / 27.4 / 27.3 ms
#define OFB_SECTOR_BATCH 1
#endif
constexpr int SECTOR_BATCH = OFB_SECTOR_BATCH;

struct SectorDev {
    int lo_bits;
    int n_classes;
    int n_con;
    int mode;  // 0 table, 1 search
    uint64_t lo_mask;
    uint64_t cls_lo, cls_hi;  // class mask restricted to the low / high half
    uint64_t plus[SECTOR_MAX_CON], minus[SECTOR_MAX_CON];
    int value[SECTOR_MAX_CON];
    const uint64_t *A;
    const uint32_t *A32;  // copy of A when every position fits 32 bits (else null)
    const uint32_t *B;
    int has_minus;        // some constraint has a non-empty minus mask
    const uint64_t *sorted;  // search mode: ascending basis states
    const uint32_t *perm;    // position of sorted[k] in the caller's order
    const uint64_t *basis;   // basis state at each position
    int64_t dim;
};

__device__ __forceinline__ bool sector_member(const SectorDev &s, uint64_t j) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < SECTOR_MAX_CON; ++k)
        if (k < s.n_con) ok = ok && (__popcll(j & s.plus[k]) - __popcll(j & s.minus[k]) == s.value[k]);
    return ok;
}

__device__ __forceinline__ int64_t sector_table_pos(const SectorDev &s, uint64_t j) {
    const uint64_t hi = j >> s.lo_bits, lo = j & s.lo_mask;
    if (s.n_classes == 1) return (int64_t)(s.A[hi] + s.B[lo]);
    const uint32_t c_lo = (uint32_t)__popcll(j & s.cls_lo), c_hi = (uint32_t)__popcll(j & s.cls_hi);
    return (int64_t)(s.A[hi * (uint32_t)s.n_classes + c_lo] + s.B[lo * (uint32_t)s.n_classes + c_hi]);
}

// position of basis state j, or -1 when j is outside the sector
__device__ __forceinline__ int64_t sector_pos(const SectorDev &s, uint64_t j) {
    if (s.mode == 0) return sector_member(s, j) ? sector_table_pos(s, j) : -1;
    int64_t lo = 0, hi = s.dim;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (s.sorted[mid] < j) lo = mid + 1; else hi = mid;
    }
    return (lo < s.dim && s.sorted[lo] == j) ? (int64_t)s.perm[lo] : -1;
}

// table mode: enumerate the members once and drop each at its position.  One thread per high
// half: halves that cannot be completed to a member (plus-only constraints: too many bits set
// already, or too few low bits left to reach the count) are skipped, so dilute sectors of 40
// qubits cost O(feasible halves * 2^lo_bits) instead of 2^40; the low-half loop reads the
// low table warp-uniformly.
__global__ void __launch_bounds__(SBLOCK) k_sector_enumerate(SectorDev s, uint64_t n_hi,
                                                             uint64_t *__restrict__ basis) {
    const uint64_t hi = (uint64_t)blockIdx.x * SBLOCK + threadIdx.x;
    if (hi >= n_hi) return;
    const uint64_t hi_part = hi << s.lo_bits;
    if (!s.has_minus) {
        for (int k = 0; k < s.n_con; ++k) {
            const int have = __popcll(hi_part & s.plus[k]);
            const int room = __popcll(s.plus[k] & s.lo_mask);
            if (have > s.value[k] || s.value[k] - have > room) return;
        }
    }
    const uint64_t n_lo = s.lo_mask + 1;
    for (uint64_t lo = 0; lo < n_lo; ++lo) {
        const uint64_t j = hi_part | lo;
        if (sector_member(s, j)) basis[sector_table_pos(s, j)] = j;
    }
}

// sigma-signed sum of `cnt` staged table entries at basis state j (two independent chains)
template <bool Z32>
__device__ __forceinline__ double sector_run_sum(const TermZC *tab, uint32_t cnt, uint64_t j) {
    double a0 = 0.0, a1 = 0.0;
    const TermZC *p = tab;
    const TermZC *end2 = tab + (cnt & ~1u);
    const uint32_t j32 = (uint32_t)j;
#pragma unroll 1
    while (p != end2) {
        const TermZC e0 = p[0], e1 = p[1];
        p += 2;
        const uint32_t s0 = Z32 ? (uint32_t)__popc(j32 & (uint32_t)e0.z) : (uint32_t)__popcll(j & e0.z);
        const uint32_t s1 = Z32 ? (uint32_t)__popc(j32 & (uint32_t)e1.z) : (uint32_t)__popcll(j & e1.z);
        a0 += __hiloint2double(__double2hiint(e0.c) + (int)(s0 << 31), __double2loint(e0.c));
        a1 += __hiloint2double(__double2hiint(e1.c) + (int)(s1 << 31), __double2loint(e1.c));
    }
    if (cnt & 1u) {
        const TermZC e0 = p[0];
        const uint32_t s0 = Z32 ? (uint32_t)__popc(j32 & (uint32_t)e0.z) : (uint32_t)__popcll(j & e0.z);
        a0 += __hiloint2double(__double2hiint(e0.c) + (int)(s0 << 31), __double2loint(e0.c));
    }
    return a0 + a1;
}

// position of j ^ x for a member j, or -1: table mode with 32-bit fast paths
template <bool Z32, bool POS32>
__device__ __forceinline__ int64_t sector_pos_fast(const SectorDev &s, uint64_t jx) {
    if (s.mode != 0) return sector_pos(s, jx);
    if (Z32 && !s.has_minus) {
        const uint32_t v = (uint32_t)jx;
        bool ok = __popc(v & (uint32_t)s.plus[0]) == s.value[0];
        if (s.n_con > 1) ok = ok && __popc(v & (uint32_t)s.plus[1]) == s.value[1];
        if (s.n_con > 2) ok = ok && sector_member(s, jx);
        if (!ok) return -1;
    } else if (!sector_member(s, jx)) {
        return -1;
    }
    if (POS32 && s.n_classes == 1) {
        if (Z32) {
            const uint32_t v = (uint32_t)jx;
            return (int64_t)(uint32_t)(s.A32[v >> s.lo_bits] + s.B[v & (uint32_t)s.lo_mask]);
        }
        return (int64_t)(uint32_t)(s.A32[jx >> s.lo_bits] + s.B[jx & s.lo_mask]);
    }
    return sector_table_pos(s, jx);
}

// out[p] = sum over segments of S(basis[p]) * in[pos(basis[p] ^ x)].  The plan's chunks (<= 16
// segments, <= 256 terms) are staged to shared memory; table reads in the term loop are
// broadcast LDS.  Per term: LDS.128, LOP, POPC, IMAD (sign fold), DADD.
template <bool Z32, bool POS32, bool HAS_IMAG, bool EXPECT>
__global__ void __launch_bounds__(SBLOCK)
k_sector_apply(SectorDev sec, const double2 *__restrict__ vin, double2 *__restrict__ vout,
               const Chunk *__restrict__ chunks, uint32_t n_chunks, const Segment *__restrict__ segs,
               const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im,
               double2 *__restrict__ partials) {
    __shared__ Segment s_segs[CHUNK_MAX_SEGS];
    __shared__ TermZC s_re[CHUNK_MAX_TERMS + 2];
    __shared__ TermZC s_im[HAS_IMAG ? CHUNK_MAX_TERMS + 2 : 1];
    const int64_t p = (int64_t)blockIdx.x * SBLOCK + threadIdx.x;
    const bool live = p < sec.dim;
    const uint64_t j = live ? sec.basis[p] : 0;
    double ar = 0.0, ai = 0.0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const Chunk ck = chunks[c];
        __syncthreads();
        {
            const uint4 *gs = (const uint4 *)(segs + ck.seg_begin);
            uint4 *ss = (uint4 *)s_segs;
            for (uint32_t u = threadIdx.x; u < ck.seg_count * 2; u += SBLOCK) ss[u] = gs[u];
            for (uint32_t u = threadIdx.x; u < ck.term_count; u += SBLOCK) {
                s_re[u] = zc_re[ck.term_begin + u];
                if (HAS_IMAG) s_im[u] = zc_im[ck.term_begin + u];
            }
        }
        __syncthreads();
        // Segments in batches of SECTOR_BATCH: all index-map lookups of a batch are issued
        // before the gathers, all gathers before the sign sums (several dependent load chains
        // table -> amplitude in flight per warp when SECTOR_BATCH > 1).
        for (uint32_t s0 = 0; s0 < ck.seg_count; s0 += SECTOR_BATCH) {
            int64_t q[SECTOR_BATCH];
            double2 v[SECTOR_BATCH];
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u) {
                q[u] = -1;
                if (live && s0 + u < ck.seg_count) {
                    const uint64_t x = Z32 ? (uint64_t)s_segs[s0 + u].x_lo : s_segs[s0 + u].x;
                    q[u] = x == 0 ? p : sector_pos_fast<Z32, POS32>(sec, j ^ x);
                }
            }
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u)
                v[u] = q[u] >= 0 ? vin[q[u]] : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u) {
                if (q[u] < 0) continue;
                const uint32_t kind_runs = s_segs[s0 + u].kind_runs;
                const uint32_t cnt = (kind_runs >> 8) & 0xffffu;
                const uint32_t tb = s_segs[s0 + u].rel_begin;
                if (!HAS_IMAG) {
                    const double sr = sector_run_sum<Z32>(s_re + tb, cnt, j);
                    ar = fma(sr, v[u].x, ar);
                    ai = fma(sr, v[u].y, ai);
                } else {
                    const uint32_t kind = kind_runs >> 30;
                    if (kind != KIND_IMAG) {
                        const double sr = sector_run_sum<Z32>(s_re + tb, cnt, j);
                        ar = fma(sr, v[u].x, ar);
                        ai = fma(sr, v[u].y, ai);
                    }
                    if (kind != KIND_REAL) {
                        const double si = sector_run_sum<Z32>(s_im + tb, cnt, j);
                        ar = fma(-si, v[u].y, ar);
                        ai = fma(si, v[u].x, ai);
                    }
                }
            }
        }
    }
    if (!EXPECT) {
        if (live) vout[p] = make_double2(ar, ai);
    } else {
        double er = 0.0, ei = 0.0;
        if (live) {
            const double2 w = vin[p];
            er = w.x * ar + w.y * ai;
            ei = w.x * ai - w.y * ar;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            er += __shfl_xor_sync(0xffffffffu, er, off);
            ei += __shfl_xor_sync(0xffffffffu, ei, off);
        }
        __shared__ double sm[2 * (SBLOCK / 32)];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            sm[2 * warp] = er;
            sm[2 * warp + 1] = ei;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tr = 0.0, ti = 0.0;
            for (int w = 0; w < SBLOCK / 32; ++w) {
                tr += sm[2 * w];
                ti += sm[2 * w + 1];
            }
            partials[blockIdx.x] = make_double2(tr, ti);
        }
    }
}

// Fast path of k_sector_apply for the sectors that matter at scale: table mode, one class,
// n <= 32 qubits, positions below 2^32, one or two "popcount == value" constraints (number
// sector / S_z + number sector).  All sector parameters live in registers, positions are
// 32-bit, and the membership test of j ^ x is two LOP/POPC/compare triples.
struct SectorFast {
    uint32_t plus0, plus1;
    int value0, value1;
    uint32_t lo_bits, lo_mask;
    const uint32_t *A32;
    const uint32_t *B;
    const uint64_t *basis;
    uint32_t dim;
};
constexpr uint32_t NO_POS = 0xffffffffu;

template <int NCON, bool HAS_IMAG, bool EXPECT>
__global__ void __launch_bounds__(SBLOCK)
k_sector_apply_fast(SectorFast sec, const double2 *__restrict__ vin, double2 *__restrict__ vout,
                    const Chunk *__restrict__ chunks, uint32_t n_chunks,
                    const Segment *__restrict__ segs, const TermZC *__restrict__ zc_re,
                    const TermZC *__restrict__ zc_im, double2 *__restrict__ partials) {
    __shared__ Segment s_segs[CHUNK_MAX_SEGS];
    __shared__ TermZC s_re[CHUNK_MAX_TERMS + 2];
    __shared__ TermZC s_im[HAS_IMAG ? CHUNK_MAX_TERMS + 2 : 1];
    const uint32_t p = blockIdx.x * SBLOCK + threadIdx.x;
    const bool live = p < sec.dim;
    const uint32_t j = live ? (uint32_t)sec.basis[p] : 0u;
    const uint32_t plus0 = sec.plus0, plus1 = sec.plus1, lo_bits = sec.lo_bits, lo_mask = sec.lo_mask;
    const int value0 = sec.value0, value1 = sec.value1;
    const uint32_t *__restrict__ tab_hi = sec.A32;
    const uint32_t *__restrict__ tab_lo = sec.B;
    double ar = 0.0, ai = 0.0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const Chunk ck = chunks[c];
        __syncthreads();
        {
            const uint4 *gs = (const uint4 *)(segs + ck.seg_begin);
            uint4 *ss = (uint4 *)s_segs;
            for (uint32_t u = threadIdx.x; u < ck.seg_count * 2; u += SBLOCK) ss[u] = gs[u];
            for (uint32_t u = threadIdx.x; u < ck.term_count; u += SBLOCK) {
                s_re[u] = zc_re[ck.term_begin + u];
                if (HAS_IMAG) s_im[u] = zc_im[ck.term_begin + u];
            }
        }
        __syncthreads();
        for (uint32_t s0 = 0; s0 < ck.seg_count; s0 += SECTOR_BATCH) {
            uint32_t q[SECTOR_BATCH];
            double2 v[SECTOR_BATCH];
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u) {
                q[u] = NO_POS;
                if (live && s0 + u < ck.seg_count) {
                    const uint32_t jx = j ^ s_segs[s0 + u].x_lo;
                    bool ok = __popc(jx & plus0) == value0;
                    if (NCON == 2) ok = ok && (__popc(jx & plus1) == value1);
                    if (ok) q[u] = tab_hi[jx >> lo_bits] + tab_lo[jx & lo_mask];
                }
            }
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u)
                v[u] = q[u] != NO_POS ? vin[q[u]] : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < SECTOR_BATCH; ++u) {
                if (q[u] == NO_POS) continue;
                const uint32_t kind_runs = s_segs[s0 + u].kind_runs;
                const uint32_t cnt = (kind_runs >> 8) & 0xffffu;
                const uint32_t tb = s_segs[s0 + u].rel_begin;
                if (!HAS_IMAG) {
                    const double sr = sector_run_sum<true>(s_re + tb, cnt, j);
                    ar = fma(sr, v[u].x, ar);
                    ai = fma(sr, v[u].y, ai);
                } else {
                    const uint32_t kind = kind_runs >> 30;
                    if (kind != KIND_IMAG) {
                        const double sr = sector_run_sum<true>(s_re + tb, cnt, j);
                        ar = fma(sr, v[u].x, ar);
                        ai = fma(sr, v[u].y, ai);
                    }
                    if (kind != KIND_REAL) {
                        const double si = sector_run_sum<true>(s_im + tb, cnt, j);
                        ar = fma(-si, v[u].y, ar);
                        ai = fma(si, v[u].x, ai);
                    }
                }
            }
        }
    }
    if (!EXPECT) {
        if (live) vout[p] = make_double2(ar, ai);
    } else {
        double er = 0.0, ei = 0.0;
        if (live) {
            const double2 w = vin[p];
            er = w.x * ar + w.y * ai;
            ei = w.x * ai - w.y * ar;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            er += __shfl_xor_sync(0xffffffffu, er, off);
            ei += __shfl_xor_sync(0xffffffffu, ei, off);
        }
        __shared__ double sm[2 * (SBLOCK / 32)];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            sm[2 * warp] = er;
            sm[2 * warp + 1] = ei;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tr = 0.0, ti = 0.0;
            for (int w = 0; w < SBLOCK / 32; ++w) {
                tr += sm[2 * w];
                ti += sm[2 * w + 1];
            }
            partials[blockIdx.x] = make_double2(tr, ti);
        }
    }
}

__global__ void __launch_bounds__(SBLOCK)
k_sector_diagonal(SectorDev sec, double *__restrict__ out, const TermZC *__restrict__ zc, uint32_t te) {
    const int64_t p = (int64_t)blockIdx.x * SBLOCK + threadIdx.x;
    if (p >= sec.dim) return;
    const uint64_t j = sec.basis[p];
    double s = 0.0;
    for (uint32_t t = 0; t < te; ++t) {
        const TermZC e = zc[t];
        const uint32_t par = (uint32_t)__popcll(j & e.z);
        s += __hiloint2double(__double2hiint(e.c) ^ (int)(par << 31), __double2loint(e.c));
    }
    out[p] = s;
}

// restrict_state / expand_state (sparse_tools.py:429-476, :507-509)
__global__ void __launch_bounds__(SBLOCK)
k_sector_gather(SectorDev sec, const double2 *__restrict__ full, double2 *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * SBLOCK + threadIdx.x;
    if (p < sec.dim) out[p] = full[sec.basis[p]];
}
__global__ void __launch_bounds__(SBLOCK)
k_sector_scatter(SectorDev sec, const double2 *__restrict__ in, double2 *__restrict__ full) {
    const int64_t p = (int64_t)blockIdx.x * SBLOCK + threadIdx.x;
    if (p < sec.dim) full[sec.basis[p]] = in[p];
}

// ---- sector CSC (scatter form: works for Pauli and for projected fermion rows) ----------
__device__ __forceinline__ void sector_group_value(const ScatterRow *__restrict__ rows, uint32_t b,
                                                   uint32_t e, uint64_t c, int n, double &sr, double &si) {
    sr = 0.0;
    si = 0.0;
    bool first = true;  // as in csc.cu: signed zeros of single-entry sums survive
    for (uint32_t k = b; k < e; ++k) {
        const ScatterRow r = rows[k];
        if ((c & r.occ_mask) == r.occ_val) {
            const double2 v = scatter_value(r, c, n);
            sr = first ? v.x : sr + v.x;
            si = first ? v.y : si + v.y;
            first = false;
        }
    }
}

// FILL = false: count the non-zero entries of column p.  FILL = true: write them in group
// order (unsorted rows) at colptr[p]...
template <bool FILL>
__global__ void __launch_bounds__(SBLOCK)
k_sector_csc_pass(SectorDev sec, const ScatterRow *__restrict__ rows,
                  const GroupDesc *__restrict__ groups, uint32_t G, int64_t *__restrict__ counts,
                  const int64_t *__restrict__ colptr, int64_t *__restrict__ tmp_rows,
                  double2 *__restrict__ tmp_vals, int n_qubits) {
    const int64_t p = (int64_t)blockIdx.x * SBLOCK + threadIdx.x;
    if (p >= sec.dim) return;
    const uint64_t c = sec.basis[p];
    int64_t at = FILL ? colptr[p] : 0;
    for (uint32_t g = 0; g < G; ++g) {
        const GroupDesc gd = groups[g];
        const int64_t row = gd.x == 0 ? p : sector_pos(sec, c ^ gd.x);
        if (row < 0) continue;
        double sr, si;
        sector_group_value(rows, gd.begin, gd.begin + (gd.count & 0x3fffffffu), c, n_qubits, sr, si);
        if (sr != 0.0 || si != 0.0) {
            if (FILL) {
                tmp_rows[at] = row;
                tmp_vals[at] = make_double2(sr, si);
            }
            ++at;
        }
    }
    if (!FILL) counts[p] = at;
}

// One warp per column: entry e goes to slot #{rows of the column smaller than row_e}
// (rows inside a column are distinct because distinct x-masks reach distinct states).
template <typename I>
__global__ void __launch_bounds__(SBLOCK)
k_sector_csc_place(int64_t dim, const int64_t *__restrict__ colptr, const int64_t *__restrict__ tmp_rows,
                   const double2 *__restrict__ tmp_vals, I *__restrict__ indices,
                   double2 *__restrict__ data) {
    const int64_t col = ((int64_t)blockIdx.x * SBLOCK + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= dim) return;
    const int64_t b = colptr[col], e = colptr[col + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
        const int64_t row = tmp_rows[k];
        int64_t rank = 0;
        for (int64_t m = b; m < e; ++m) rank += tmp_rows[m] < row ? 1 : 0;
        indices[b + rank] = (I)row;
        data[b + rank] = tmp_vals[k];
    }
}

template <typename I>
__global__ void k_sector_convert_ptr(const int64_t *__restrict__ in, int64_t n, I *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (I)in[i];
}

// exclusive scan of per-column counts (single block; the count pass dominates)
__global__ void __launch_bounds__(1024) k_sector_scan(const int64_t *__restrict__ counts, int64_t n,
                                                      int64_t *__restrict__ colptr) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n ? counts[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int64_t t = warp_tot[lane];
            int64_t ti = t;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t o = __shfl_up_sync(0xffffffffu, ti, off);
                if (lane >= off) ti += o;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const int64_t c0 = carry;
        if (i < n) colptr[i] = c0 + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + warp_tot[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) colptr[n] = carry;
}

}  // namespace ofb

using namespace ofb;

struct ofb_sector {
    int n_qubits = 0;
    int device = 0;
    int64_t dim = 0;
    SectorDev dev{};
    uint64_t *d_A = nullptr;
    uint32_t *d_A32 = nullptr;
    uint32_t *d_B = nullptr;
    uint64_t *d_sorted = nullptr;
    uint32_t *d_perm = nullptr;
    uint64_t *d_basis = nullptr;
    // CSC state between count and fill
    int64_t *d_colptr = nullptr;  // dim + 1, then counts[dim]
    int64_t csc_nnz = -1;
    // staging for *_host calls
    void *d_stage_in = nullptr, *d_stage_out = nullptr;
    size_t stage_bytes = 0;
    double *d_partials = nullptr;
    size_t partials_bytes = 0;
    double *h_pinned = nullptr;
};

static inline unsigned sector_grid(int64_t dim) { return (unsigned)((dim + SBLOCK - 1) / SBLOCK); }

template <bool EXPECT>
static int sector_launch_apply(ofb_plan *p, ofb_sector *s, const double2 *in, double2 *out,
                               double2 *partials, cudaStream_t st) {
    const unsigned grid = sector_grid(s->dim);
    const uint32_t n_chunks = (uint32_t)p->n_chunks;
    const bool z32 = p->n_qubits <= 32, pos32 = s->dev.A32 != nullptr;
    const SectorDev &d = s->dev;
    if (z32 && pos32 && d.mode == 0 && d.n_classes == 1 && !d.has_minus && d.n_con <= 2 &&
        s->dim < 0xffffffffll && !getenv("OFB_SECTOR_GENERIC")) {
        SectorFast f;
        f.plus0 = (uint32_t)d.plus[0];
        f.value0 = d.value[0];
        f.plus1 = d.n_con > 1 ? (uint32_t)d.plus[1] : 0u;
        f.value1 = d.n_con > 1 ? d.value[1] : 0;
        f.lo_bits = (uint32_t)d.lo_bits;
        f.lo_mask = (uint32_t)d.lo_mask;
        f.A32 = d.A32;
        f.B = d.B;
        f.basis = d.basis;
        f.dim = (uint32_t)s->dim;
#define OFB_SFAST(NN, II)                                                                        \
    k_sector_apply_fast<NN, II, EXPECT><<<grid, SBLOCK, 0, st>>>(f, in, out, p->d_chunks, n_chunks, \
                                                                 p->d_segs, p->d_zc, p->d_zc_im,  \
                                                                 partials)
        if (d.n_con == 1) { if (p->has_imag) OFB_SFAST(1, true); else OFB_SFAST(1, false); }
        else              { if (p->has_imag) OFB_SFAST(2, true); else OFB_SFAST(2, false); }
#undef OFB_SFAST
        OFB_LAUNCH_CHECK();
        return OFB_OK;
    }
#define OFB_SGO(ZZ, PP, II)                                                                      \
    k_sector_apply<ZZ, PP, II, EXPECT><<<grid, SBLOCK, 0, st>>>(s->dev, in, out, p->d_chunks,    \
                                                                n_chunks, p->d_segs, p->d_zc,    \
                                                                p->d_zc_im, partials)
    if (z32) {
        if (pos32) { if (p->has_imag) OFB_SGO(true, true, true); else OFB_SGO(true, true, false); }
        else       { if (p->has_imag) OFB_SGO(true, false, true); else OFB_SGO(true, false, false); }
    } else {
        if (pos32) { if (p->has_imag) OFB_SGO(false, true, true); else OFB_SGO(false, true, false); }
        else       { if (p->has_imag) OFB_SGO(false, false, true); else OFB_SGO(false, false, false); }
    }
#undef OFB_SGO
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

static int sector_common_checks(const ofb_plan *p, const ofb_sector *s) {
    if (!p || !s) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->n_qubits != s->n_qubits) return fail(OFB_ERR_INVALID, "plan and sector differ in n_qubits");
    if (p->device != s->device) return fail(OFB_ERR_INVALID, "plan and sector live on different devices");
    return OFB_OK;
}

extern "C" {

int ofb_sector_create_tables(int n_qubits, int lo_bits, int n_classes, uint64_t class_mask,
                             const uint64_t *table_hi, const uint32_t *table_lo, int n_constraints,
                             const uint64_t *plus_mask, const uint64_t *minus_mask,
                             const int32_t *value, int64_t dim, int device, ofb_sector **sector) {
    if (!sector) return fail(OFB_ERR_INVALID, "sector is NULL");
    *sector = nullptr;
    ofb_sector *s = nullptr;
    OFB_TRY
    if (n_qubits < 1 || n_qubits > 40) return fail(OFB_ERR_INVALID, "n_qubits out of range [1, 40]");
    if (lo_bits < 0 || lo_bits > n_qubits || lo_bits > 26 || n_qubits - lo_bits > 26)
        return fail(OFB_ERR_INVALID, "bad lo_bits");
    if (n_classes < 1 || n_classes > 64) return fail(OFB_ERR_INVALID, "bad n_classes");
    if (n_constraints < 1 || n_constraints > SECTOR_MAX_CON)
        return fail(OFB_ERR_INVALID, "need 1..4 membership constraints");
    if (!table_hi || !table_lo || !plus_mask || !minus_mask || !value)
        return fail(OFB_ERR_INVALID, "NULL table");
    if (dim < 0 || dim > (1ll << 33)) return fail(OFB_ERR_INVALID, "sector dimension out of range");
    OFB_CUDA(cudaSetDevice(device));
    s = new (std::nothrow) ofb_sector();
    if (!s) return fail(OFB_ERR_NOMEM, "host allocation failed");
    s->n_qubits = n_qubits;
    s->device = device;
    s->dim = dim;
    const size_t n_hi = ((size_t)1 << (n_qubits - lo_bits)) * n_classes;
    const size_t n_lo = ((size_t)1 << lo_bits) * n_classes;
    cudaError_t e;
    if ((e = cudaMalloc(&s->d_A, n_hi * 8)) != cudaSuccess || (e = cudaMalloc(&s->d_B, n_lo * 4)) != cudaSuccess ||
        (e = cudaMalloc(&s->d_basis, std::max<int64_t>(dim, 1) * 8)) != cudaSuccess ||
        (e = cudaMallocHost(&s->h_pinned, 64 * sizeof(double))) != cudaSuccess ||
        (e = cudaMemcpy(s->d_A, table_hi, n_hi * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(s->d_B, table_lo, n_lo * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        int rc = cuda_fail(e, "sector tables", __FILE__, __LINE__);
        ofb_sector_destroy(s);
        return rc;
    }
    SectorDev &d = s->dev;
    d.lo_bits = lo_bits;
    d.n_classes = n_classes;
    d.n_con = n_constraints;
    d.mode = 0;
    d.lo_mask = lo_bits >= 64 ? ~0ull : ((1ull << lo_bits) - 1);
    d.cls_lo = class_mask & d.lo_mask;
    d.cls_hi = class_mask & ~d.lo_mask;
    for (int k = 0; k < n_constraints; ++k) {
        d.plus[k] = plus_mask[k];
        d.minus[k] = minus_mask[k];
        d.value[k] = value[k];
    }
    d.A = s->d_A;
    d.B = s->d_B;
    d.basis = s->d_basis;
    d.dim = dim;
    d.has_minus = 0;
    for (int k = 0; k < n_constraints; ++k) d.has_minus |= minus_mask[k] != 0;
    d.A32 = nullptr;
    if (dim < (1ll << 32)) {
        std::vector<uint32_t> a32(n_hi);
        for (size_t k = 0; k < n_hi; ++k) a32[k] = (uint32_t)table_hi[k];
        if ((e = cudaMalloc(&s->d_A32, n_hi * 4)) != cudaSuccess ||
            (e = cudaMemcpy(s->d_A32, a32.data(), n_hi * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            int rc = cuda_fail(e, "sector tables", __FILE__, __LINE__);
            ofb_sector_destroy(s);
            return rc;
        }
        d.A32 = s->d_A32;
    }
    if (dim > 0) {
        const uint64_t n_hi_states = 1ull << (n_qubits - lo_bits);
        const unsigned grid = (unsigned)((n_hi_states + SBLOCK - 1) / SBLOCK);
        // poison first so that a wrong table (a position never written) is caught by the tests
        cudaMemset(s->d_basis, 0xff, dim * 8);
        k_sector_enumerate<<<grid, SBLOCK>>>(d, n_hi_states, s->d_basis);
        g_launches.fetch_add(1);
        if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
            int rc = cuda_fail(e, "k_sector_enumerate", __FILE__, __LINE__);
            ofb_sector_destroy(s);
            return rc;
        }
    }
    *sector = s;
    return OFB_OK;
    OFB_CATCH(ofb_sector_destroy(s))
}

int ofb_sector_create_list(int n_qubits, int64_t dim, const uint64_t *h_basis, int device,
                           ofb_sector **sector) {
    if (!sector) return fail(OFB_ERR_INVALID, "sector is NULL");
    *sector = nullptr;
    ofb_sector *s = nullptr;
    OFB_TRY
    if (n_qubits < 0 || n_qubits > 40) return fail(OFB_ERR_INVALID, "n_qubits out of range [0, 40]");
    if (dim < 0 || dim >= (1ll << 32)) return fail(OFB_ERR_INVALID, "basis list too long");
    if (dim > 0 && !h_basis) return fail(OFB_ERR_INVALID, "NULL basis");
    const uint64_t full = (1ull << n_qubits) - 1;
    std::vector<uint32_t> perm((size_t)dim);
    for (int64_t k = 0; k < dim; ++k) {
        if (h_basis[k] & ~full) return fail(OFB_ERR_INVALID, "basis state outside n_qubits");
        perm[k] = (uint32_t)k;
    }
    std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return h_basis[a] < h_basis[b]; });
    std::vector<uint64_t> sorted((size_t)dim);
    for (int64_t k = 0; k < dim; ++k) {
        sorted[k] = h_basis[perm[k]];
        if (k > 0 && sorted[k] == sorted[k - 1]) return fail(OFB_ERR_INVALID, "duplicate basis state");
    }
    OFB_CUDA(cudaSetDevice(device));
    s = new (std::nothrow) ofb_sector();
    if (!s) return fail(OFB_ERR_NOMEM, "host allocation failed");
    s->n_qubits = n_qubits;
    s->device = device;
    s->dim = dim;
    const size_t m = (size_t)std::max<int64_t>(dim, 1);
    cudaError_t e;
    if ((e = cudaMalloc(&s->d_sorted, m * 8)) != cudaSuccess || (e = cudaMalloc(&s->d_perm, m * 4)) != cudaSuccess ||
        (e = cudaMalloc(&s->d_basis, m * 8)) != cudaSuccess ||
        (e = cudaMallocHost(&s->h_pinned, 64 * sizeof(double))) != cudaSuccess ||
        (dim && (e = cudaMemcpy(s->d_sorted, sorted.data(), dim * 8, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (dim && (e = cudaMemcpy(s->d_perm, perm.data(), dim * 4, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (dim && (e = cudaMemcpy(s->d_basis, h_basis, dim * 8, cudaMemcpyHostToDevice)) != cudaSuccess)) {
        int rc = cuda_fail(e, "sector list", __FILE__, __LINE__);
        ofb_sector_destroy(s);
        return rc;
    }
    SectorDev &d = s->dev;
    d.mode = 1;
    d.n_classes = 1;
    d.n_con = 0;
    d.sorted = s->d_sorted;
    d.perm = s->d_perm;
    d.basis = s->d_basis;
    d.dim = dim;
    *sector = s;
    return OFB_OK;
    OFB_CATCH(ofb_sector_destroy(s))
}

int ofb_sector_destroy(ofb_sector *s) {
    if (!s) return OFB_OK;
    cudaSetDevice(s->device);
    cudaFree(s->d_A);
    cudaFree(s->d_A32);
    cudaFree(s->d_B);
    cudaFree(s->d_sorted);
    cudaFree(s->d_perm);
    cudaFree(s->d_basis);
    cudaFree(s->d_colptr);
    cudaFree(s->d_stage_in);
    cudaFree(s->d_stage_out);
    cudaFree(s->d_partials);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    delete s;
    return OFB_OK;
}

int ofb_sector_info(const ofb_sector *s, int *n_qubits, int64_t *dim, int *mode) {
    if (!s) return fail(OFB_ERR_INVALID, "NULL argument");
    if (n_qubits) *n_qubits = s->n_qubits;
    if (dim) *dim = s->dim;
    if (mode) *mode = s->dev.mode;
    return OFB_OK;
}

int ofb_sector_basis(ofb_sector *s, uint64_t *h_basis) {
    if (!s || !h_basis) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(s->device));
    if (s->dim) OFB_CUDA(cudaMemcpy(h_basis, s->d_basis, s->dim * 8, cudaMemcpyDeviceToHost));
    return OFB_OK;
}

int ofb_sector_apply(ofb_plan *p, ofb_sector *s, const void *d_in, void *d_out, int ncols, int64_t ld,
                     void *stream) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!d_in || !d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    if (ncols < 0 || (ncols > 1 && ld < s->dim)) return fail(OFB_ERR_INVALID, "bad ncols/ld");
    if (d_in == d_out) return fail(OFB_ERR_INVALID, "d_in and d_out must not alias");
    OFB_CUDA(cudaSetDevice(p->device));
    if (s->dim == 0) return OFB_OK;
    for (int c = 0; c < ncols; ++c) {
        const double2 *in = (const double2 *)d_in + (size_t)c * ld;
        double2 *out = (double2 *)d_out + (size_t)c * ld;
        if ((rc = sector_launch_apply<false>(p, s, in, out, nullptr, (cudaStream_t)stream))) return rc;
    }
    return OFB_OK;
}

static int sector_stage(ofb_sector *s, size_t bytes) {
    if (s->stage_bytes >= bytes) return OFB_OK;
    cudaFree(s->d_stage_in);
    cudaFree(s->d_stage_out);
    s->d_stage_in = s->d_stage_out = nullptr;
    s->stage_bytes = 0;
    OFB_CUDA(cudaMalloc(&s->d_stage_in, bytes));
    OFB_CUDA(cudaMalloc(&s->d_stage_out, bytes));
    s->stage_bytes = bytes;
    return OFB_OK;
}

int ofb_sector_apply_host(ofb_plan *p, ofb_sector *s, const void *h_in, void *h_out, int ncols) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!h_in || !h_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (ncols < 0) return fail(OFB_ERR_INVALID, "bad ncols");
    OFB_CUDA(cudaSetDevice(p->device));
    if (s->dim == 0) return OFB_OK;
    const size_t col_bytes = (size_t)s->dim * 16;
    if ((rc = sector_stage(s, col_bytes))) return rc;
    for (int c = 0; c < ncols; ++c) {
        OFB_CUDA(cudaMemcpyAsync(s->d_stage_in, (const char *)h_in + c * col_bytes, col_bytes,
                                 cudaMemcpyHostToDevice, 0));
        if ((rc = ofb_sector_apply(p, s, s->d_stage_in, s->d_stage_out, 1, s->dim, nullptr))) return rc;
        OFB_CUDA(cudaMemcpyAsync((char *)h_out + c * col_bytes, s->d_stage_out, col_bytes,
                                 cudaMemcpyDeviceToHost, 0));
    }
    OFB_CUDA(cudaStreamSynchronize(0));
    return OFB_OK;
}

int ofb_sector_expectation(ofb_plan *p, ofb_sector *s, const void *d_state, double *h_out_ri,
                           void *stream) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!d_state || !h_out_ri) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    h_out_ri[0] = h_out_ri[1] = 0.0;
    if (s->dim == 0 || p->n_groups == 0) return OFB_OK;
    const unsigned grid = sector_grid(s->dim);
    const size_t need = ((size_t)grid * 2 + 2) * sizeof(double);
    if (s->partials_bytes < need) {
        cudaFree(s->d_partials);
        s->d_partials = nullptr;
        s->partials_bytes = 0;
        OFB_CUDA(cudaMalloc(&s->d_partials, need));
        s->partials_bytes = need;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = sector_launch_apply<true>(p, s, (const double2 *)d_state, nullptr,
                                        (double2 *)s->d_partials, st)))
        return rc;
    double *d_res = s->d_partials + 2 * (size_t)grid;
    if ((rc = reduce_partials(s->d_partials, (int64_t)grid, 2, d_res, st))) return rc;
    OFB_CUDA(cudaMemcpyAsync(s->h_pinned, d_res, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFB_CUDA(cudaStreamSynchronize(st));
    h_out_ri[0] = s->h_pinned[0];
    h_out_ri[1] = s->h_pinned[1];
    return OFB_OK;
}

int ofb_sector_diagonal(ofb_plan *p, ofb_sector *s, double *d_out, void *stream) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    if (s->dim == 0) return OFB_OK;
    k_sector_diagonal<<<sector_grid(s->dim), SBLOCK, 0, (cudaStream_t)stream>>>(
        s->dev, d_out, p->d_zc, (uint32_t)p->n_diag_terms);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_sector_gather(ofb_sector *s, const void *d_full, void *d_sector, void *stream) {
    if (!s || !d_full || !d_sector) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(s->device));
    if (s->dim == 0) return OFB_OK;
    k_sector_gather<<<sector_grid(s->dim), SBLOCK, 0, (cudaStream_t)stream>>>(
        s->dev, (const double2 *)d_full, (double2 *)d_sector);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_sector_scatter(ofb_sector *s, const void *d_sector, void *d_full, void *stream) {
    if (!s || !d_sector || !d_full) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(s->device));
    OFB_CUDA(cudaMemsetAsync(d_full, 0, ((size_t)1 << s->n_qubits) * 16, (cudaStream_t)stream));
    if (s->dim == 0) return OFB_OK;
    k_sector_scatter<<<sector_grid(s->dim), SBLOCK, 0, (cudaStream_t)stream>>>(
        s->dev, (const double2 *)d_sector, (double2 *)d_full);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

static int sector_ensure_groups(ofb_plan *p) {
    if (p->d_groups) return OFB_OK;
    OFB_TRY
    std::vector<GroupDesc> groups(p->n_groups);
    for (int64_t g = 0; g < p->n_groups; ++g)
        groups[g] = GroupDesc{p->h_group_x[g], p->h_group_begin[g],
                              p->h_group_begin[g + 1] - p->h_group_begin[g]};
    OFB_CUDA(cudaMalloc(&p->d_groups, std::max<size_t>(groups.size(), 1) * sizeof(GroupDesc)));
    if (!groups.empty())
        OFB_CUDA(cudaMemcpy(p->d_groups, groups.data(), groups.size() * sizeof(GroupDesc),
                            cudaMemcpyHostToDevice));
    return OFB_OK;
    OFB_CATCH((void)0)
}

int ofb_sector_csc_count(ofb_plan *p, ofb_sector *s, int64_t *nnz, int *index_bits, void *stream) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!nnz) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(p->device));
    if ((rc = sector_ensure_groups(p))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaFree(s->d_colptr);
    s->d_colptr = nullptr;
    s->csc_nnz = -1;
    const int64_t dim = s->dim;
    OFB_CUDA(cudaMalloc(&s->d_colptr, (2 * dim + 2) * sizeof(int64_t)));
    int64_t *colptr = s->d_colptr, *counts = s->d_colptr + dim + 1;
    if (dim > 0) {
        k_sector_csc_pass<false><<<sector_grid(dim), SBLOCK, 0, st>>>(
            s->dev, p->d_rows, p->d_groups, (uint32_t)p->n_groups, counts, nullptr, nullptr, nullptr,
            p->n_qubits);
        OFB_LAUNCH_CHECK();
    }
    k_sector_scan<<<1, 1024, 0, st>>>(counts, dim, colptr);
    OFB_LAUNCH_CHECK();
    OFB_CUDA(cudaMemcpyAsync(s->h_pinned, colptr + dim, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    OFB_CUDA(cudaStreamSynchronize(st));
    int64_t total;
    memcpy(&total, s->h_pinned, sizeof(int64_t));
    s->csc_nnz = total;
    *nnz = total;
    if (index_bits) *index_bits = (std::max<int64_t>(total, dim) < (1ll << 31)) ? 32 : 64;
    return OFB_OK;
}

int ofb_sector_csc_fill_host(ofb_plan *p, ofb_sector *s, void *h_indptr, void *h_indices, void *h_data,
                             int index_bits) {
    int rc = sector_common_checks(p, s);
    if (rc) return rc;
    if (!h_indptr) return fail(OFB_ERR_INVALID, "NULL argument");
    if (index_bits != 32 && index_bits != 64) return fail(OFB_ERR_INVALID, "index_bits must be 32 or 64");
    if (s->csc_nnz < 0 || !s->d_colptr) return fail(OFB_ERR_STATE, "ofb_sector_csc_count must run first");
    OFB_CUDA(cudaSetDevice(p->device));
    const int64_t dim = s->dim;
    const size_t nnz = (size_t)s->csc_nnz, ib = index_bits / 8;
    if (nnz && (!h_indices || !h_data)) return fail(OFB_ERR_INVALID, "NULL output");
    int64_t *tmp_rows = nullptr;
    double2 *tmp_vals = nullptr;
    void *d_indptr = nullptr, *d_indices = nullptr, *d_data = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&tmp_rows, std::max<size_t>(nnz, 1) * 8)) != cudaSuccess ||
        (e = cudaMalloc(&tmp_vals, std::max<size_t>(nnz, 1) * 16)) != cudaSuccess ||
        (e = cudaMalloc(&d_indptr, (size_t)(dim + 1) * ib)) != cudaSuccess ||
        (e = cudaMalloc(&d_indices, std::max<size_t>(nnz, 1) * ib)) != cudaSuccess ||
        (e = cudaMalloc(&d_data, std::max<size_t>(nnz, 1) * 16)) != cudaSuccess) {
        rc = cuda_fail(e, "cudaMalloc(sector csc output)", __FILE__, __LINE__);
    }
    if (!rc) {
        if (dim > 0 && nnz > 0) {
            k_sector_csc_pass<true><<<sector_grid(dim), SBLOCK>>>(s->dev, p->d_rows, p->d_groups,
                                                                  (uint32_t)p->n_groups, nullptr,
                                                                  s->d_colptr, tmp_rows, tmp_vals,
                                                                  p->n_qubits);
            g_launches.fetch_add(1);
            const unsigned wgrid = (unsigned)(((size_t)dim * 32 + SBLOCK - 1) / SBLOCK);
            if (index_bits == 32)
                k_sector_csc_place<int32_t><<<wgrid, SBLOCK>>>(dim, s->d_colptr, tmp_rows, tmp_vals,
                                                               (int32_t *)d_indices, (double2 *)d_data);
            else
                k_sector_csc_place<int64_t><<<wgrid, SBLOCK>>>(dim, s->d_colptr, tmp_rows, tmp_vals,
                                                               (int64_t *)d_indices, (double2 *)d_data);
            g_launches.fetch_add(1);
        }
        const unsigned pgrid = (unsigned)((dim + 1 + 255) / 256);
        if (index_bits == 32)
            k_sector_convert_ptr<int32_t><<<pgrid, 256>>>(s->d_colptr, dim + 1, (int32_t *)d_indptr);
        else
            k_sector_convert_ptr<int64_t><<<pgrid, 256>>>(s->d_colptr, dim + 1, (int64_t *)d_indptr);
        g_launches.fetch_add(1);
        if ((e = cudaGetLastError()) != cudaSuccess ||
            (e = cudaMemcpy(h_indptr, d_indptr, (size_t)(dim + 1) * ib, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (nnz && (e = cudaMemcpy(h_indices, d_indices, nnz * ib, cudaMemcpyDeviceToHost)) != cudaSuccess) ||
            (nnz && (e = cudaMemcpy(h_data, d_data, nnz * 16, cudaMemcpyDeviceToHost)) != cudaSuccess))
            rc = cuda_fail(e, "sector csc fill", __FILE__, __LINE__);
    }
    cudaFree(tmp_rows);
    cudaFree(tmp_vals);
    cudaFree(d_indptr);
    cudaFree(d_indices);
    cudaFree(d_data);
    cudaFree(s->d_colptr);
    s->d_colptr = nullptr;
    s->csc_nnz = -1;
    return rc;
}

}  // extern "C"
