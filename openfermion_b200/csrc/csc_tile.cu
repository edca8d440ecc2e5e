// K4 v2: tiled CSC passes for Pauli plans (qubit_operator_sparse, linalg/sparse_tools.py:136-194).
//
// The first builder (csc.cu) gives every column its own thread, which re-evaluates all T sign
// terms per column in both passes from global memory.  Here the count and fill passes share the
// geometry and the sign algebra of the matrix-free kernel K1 (apply.cu): a CTA of 64 threads owns
// a tile of 1 024 columns, thread `tid` owns the 16 columns c_r = tile + 64 r + tid, and the entry
// of group g in column c is the gather-form coefficient at its ROW,
//     H[c ^ x_g, c] = S_g(c ^ x_g),   S_g(j) = sum_t k_t (-1)^popc(j & z_t),   k_t = coeff_t (-i)^y_t.
// The rows j_r = (tile + tid) ^ x ^ (r << 6) of a thread differ only in the class bits, so one
// class-run sum A (LDS.128 + LOP3 + POPC + IMAD + DADD per term, tables staged in shared memory
// with cp.async) serves 16 columns and the 16 values are +-A combinations: ~6x fewer instructions
// per (group, column) than one thread per column.  Rows inside a column ascend through the same
// sibling-subtree ranks as before (plan.cu build_sibling_table), with the rank split into a
// CTA-uniform part (bits >= 10, one warp reduction), a lane part (bits 0..5) and 16 class parts.
//
// Summation order inside a group is the class-sorted order of the gather tables, not the caller's
// term order: identical to the reference whenever the sums are exact (dyadic coefficients) and
// inside the documented residue contract otherwise; ofb_plan_set_csc_mode(1)
// still replays scipy's order exactly.  Signed zeros: a part that no term of the group can make
// non-zero is +0 unless every term of the group needs the kron-chain replay (strings acting on
// the last qubit); those few groups take the per-row path of csc.cu (scatter_value).
#include <algorithm>

#include "ofb_internal.cuh"

namespace ofb {

constexpr int TB = 64;                 // threads per CTA (== 1 << CLASS_SHIFT)
constexpr int TR = 1 << CLASS_BITS;    // columns per thread
constexpr int TTILE = TB * TR;
constexpr int TTILE_BITS = CLASS_SHIFT + CLASS_BITS;
static_assert(CLASS_BITS == 4 && CLASS_SHIFT == 6, "csc_tile.cu assumes the 64 x 16 geometry");

struct CscStage {
    Segment segs[CHUNK_MAX_SEGS];
    TermZC re[CHUNK_MAX_TERMS + 8];
    TermZC im[CHUNK_MAX_TERMS + 8];
};

__device__ __forceinline__ void csc_cp16(void *smem, const void *gmem) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}

__device__ __forceinline__ void csc_stage(CscStage &b, const Chunk ck, const Segment *__restrict__ segs,
                                          const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im,
                                          bool has_imag) {
    const char *gs = (const char *)(segs + ck.seg_begin);
    char *ss = (char *)b.segs;
    for (uint32_t u = threadIdx.x; u < ck.seg_count * 2; u += TB) csc_cp16(ss + u * 16, gs + u * 16);
    for (uint32_t u = threadIdx.x; u < ck.term_count; u += TB) {
        csc_cp16(&b.re[u], zc_re + ck.term_begin + u);
        if (has_imag) csc_cp16(&b.im[u], zc_im + ck.term_begin + u);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__device__ __forceinline__ double csc_signed(const TermZC &e, uint32_t j) {
    return __hiloint2double(__double2hiint(e.c) + (int)((uint32_t)__popc(j & (uint32_t)e.z) << 31),
                            __double2loint(e.c));
}

// sum of one class run at row bits j (two accumulators, as K1's run_sum)
__device__ __forceinline__ double csc_run_sum(const TermZC *tab, uint32_t cnt, uint32_t j) {
    double a0 = 0.0, a1 = 0.0;
    uint32_t t = 0;
    for (; t + 1 < cnt; t += 2) {
        a0 += csc_signed(tab[t], j);
        a1 += csc_signed(tab[t + 1], j);
    }
    if (t < cnt) a0 += csc_signed(tab[t], j);
    return a0 + a1;
}

// v_r += (-1)^popc(r & w) a
template <int W>
__device__ __forceinline__ void csc_spread(double a, double (&v)[TR]) {
#pragma unroll
    for (int r = 0; r < TR; ++r) v[r] += (__builtin_popcount(r & W) & 1) ? -a : a;
}
__device__ __forceinline__ void csc_spread_dyn(uint32_t w, double a, double (&v)[TR]) {
    switch (w) {
        case 0: csc_spread<0>(a, v); break;
        case 1: csc_spread<1>(a, v); break;
        case 2: csc_spread<2>(a, v); break;
        case 3: csc_spread<3>(a, v); break;
        case 4: csc_spread<4>(a, v); break;
        case 5: csc_spread<5>(a, v); break;
        case 6: csc_spread<6>(a, v); break;
        case 7: csc_spread<7>(a, v); break;
        case 8: csc_spread<8>(a, v); break;
        case 9: csc_spread<9>(a, v); break;
        case 10: csc_spread<10>(a, v); break;
        case 11: csc_spread<11>(a, v); break;
        case 12: csc_spread<12>(a, v); break;
        case 13: csc_spread<13>(a, v); break;
        case 14: csc_spread<14>(a, v); break;
        default: csc_spread<15>(a, v); break;
    }
}

// per-row evaluation of a group (term order, signed zeros replayed): the rare groups whose zero
// part can be -0
__device__ __noinline__ double2 csc_group_rows(const ScatterRow *__restrict__ rows, uint32_t b, uint32_t e,
                                               uint64_t c, int n) {
    double sr = 0.0, si = 0.0;
    bool first = true;
    for (uint32_t k = b; k < e; ++k) {
        if (rows[k].cr == 0.0 && rows[k].ci == 0.0) continue;  // no triplet in the reference
        const double2 v = scatter_value(rows[k], c, n);
        sr = first ? v.x : sr + v.x;
        si = first ? v.y : si + v.y;
        first = false;
    }
    return make_double2(sr, si);
}

template <bool FILL, bool HAS_IMAG, typename I>
__global__ void __launch_bounds__(TB, 6)
k_csc_tile(const Chunk *__restrict__ chunks, uint32_t n_chunks, const Segment *__restrict__ segs,
           const uint32_t *__restrict__ seg_group, const TermZC *__restrict__ zc_re,
           const TermZC *__restrict__ zc_im, const uint32_t *__restrict__ sib,
           const uint8_t *__restrict__ group_slow, const ScatterRow *__restrict__ rows,
           const GroupDesc *__restrict__ groups, int n, uint64_t dim, uint32_t words,
           uint32_t *__restrict__ bitmap, uint32_t *__restrict__ prefix, int64_t *__restrict__ colcount,
           const int64_t *__restrict__ colptr, I *__restrict__ indices, double2 *__restrict__ data) {
    __shared__ CscStage sbuf[2];
    __shared__ uint32_t s_sib[2][CHUNK_MAX_SEGS][32];
    __shared__ uint32_t s_group[2][CHUNK_MAX_SEGS];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t c0 = blockIdx.x * (uint32_t)TTILE + tid;  // column of r = 0

    double vr[TR], vi[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) vr[r] = vi[r] = 0.0;

    auto stage_aux = [&](int buf, const Chunk ck) {  // group ids and sibling rows of the chunk's segments
        for (uint32_t q = 0; q < ck.seg_count; ++q) {
            const uint32_t g = seg_group[ck.seg_begin + q];
            if (tid == 0) s_group[buf][q] = g;
            if (tid < 32) s_sib[buf][q][tid] = tid < (uint32_t)n ? sib[(size_t)g * n + tid] : 0u;
        }
    };

    Chunk cur = chunks[0];
    Chunk nxt = n_chunks > 1 ? chunks[1] : cur;
    csc_stage(sbuf[0], cur, segs, zc_re, zc_im, HAS_IMAG);
    stage_aux(0, cur);
    for (uint32_t c = 0; c < n_chunks; ++c) {
        Chunk after = nxt;
        if (c + 1 < n_chunks) {
            csc_stage(sbuf[(c + 1) & 1], nxt, segs, zc_re, zc_im, HAS_IMAG);
            stage_aux((c + 1) & 1, nxt);
            if (c + 2 < n_chunks) after = chunks[c + 2];
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const CscStage &b = sbuf[c & 1];
        for (uint32_t q = 0; q < cur.seg_count; ++q) {
            const Segment &sd = b.segs[q];
            const uint32_t g = s_group[c & 1][q];
            const uint32_t runs = sd.kind_runs & 0xfu, kind = sd.kind_runs >> 30;
            const bool last = (sd.kind_runs >> 4) & 1u;  // last segment of its group
            const uint32_t x = (uint32_t)sd.x;
            const uint32_t jb = c0 ^ x;                  // row of r = 0 (class bits of x included)
            const bool slow = group_slow[g] != 0;
            if (!slow) {
                const TermZC *tre = b.re + sd.rel_begin, *tim = b.im + sd.rel_begin;
                uint64_t cnts = sd.cnts;
                uint32_t ids = sd.ids;
#pragma unroll 1
                for (uint32_t k = 0; k < runs; ++k) {
                    const uint32_t cnt = (uint32_t)cnts & 0xffu;
                    if (!HAS_IMAG || kind != KIND_IMAG) csc_spread_dyn(ids & 0xfu, csc_run_sum(tre, cnt, jb), vr);
                    if (HAS_IMAG && kind != KIND_REAL) csc_spread_dyn(ids & 0xfu, csc_run_sum(tim, cnt, jb), vi);
                    tre += cnt;
                    tim += cnt;
                    cnts >>= 8;
                    ids >>= 4;
                }
            }
            if (!last) continue;
            if (slow) {
                const GroupDesc gd = groups[g];
#pragma unroll
                for (int r = 0; r < TR; ++r) {  // unrolled: vr / vi must stay in registers
                    const double2 v = csc_group_rows(rows, gd.begin, gd.begin + (gd.count & 0x3fffffffu),
                                                     (uint64_t)(c0 + r * TB), n);
                    vr[r] = v.x;
                    vi[r] = v.y;
                }
            }
            // ranks of the rows j_r = jb ^ (r << 6) among the rows of their columns
            const uint32_t *sr_ = s_sib[c & 1][q];
            uint32_t part = (lane + TTILE_BITS < 32u && ((jb >> (lane + TTILE_BITS)) & 1u)) ? sr_[lane + TTILE_BITS] : 0u;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            uint32_t base = part;
#pragma unroll
            for (int bit = 0; bit < CLASS_SHIFT; ++bit) base += ((jb >> bit) & 1u) ? sr_[bit] : 0u;
            int32_t delta[CLASS_BITS];
#pragma unroll
            for (int k = 0; k < CLASS_BITS; ++k) {
                const uint32_t sv = sr_[CLASS_SHIFT + k];
                const bool set = (jb >> (CLASS_SHIFT + k)) & 1u;
                base += set ? sv : 0u;
                delta[k] = set ? -(int32_t)sv : (int32_t)sv;  // toggling bit k of r toggles that row bit
            }
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const double re = vr[r], im = vi[r];
                vr[r] = 0.0;
                vi[r] = 0.0;
                if (re == 0.0 && im == 0.0) continue;
                int32_t rank = (int32_t)base;
#pragma unroll
                for (int k = 0; k < CLASS_BITS; ++k) rank += ((r >> k) & 1) ? delta[k] : 0;
                const uint64_t col = (uint64_t)(c0 + r * TB);
                const size_t at = (size_t)((uint32_t)rank >> 5) * dim + col;
                if (!FILL) {
                    bitmap[at] |= 1u << (rank & 31);
                } else {
                    const uint32_t below = bitmap[at] & ((1u << (rank & 31)) - 1u);
                    const int64_t pos = colptr[col] + prefix[at] + __popc(below);
                    indices[pos] = (I)(col ^ (uint64_t)x);
                    data[pos] = make_double2(re, im);
                }
            }
        }
        __syncthreads();
        cur = nxt;
        nxt = after;
    }
    if (!FILL) {
#pragma unroll 1
        for (int r = 0; r < TR; ++r) {
            const uint64_t col = (uint64_t)(c0 + r * TB);
            uint32_t run = 0;
            for (uint32_t w = 0; w < words; ++w) {
                prefix[(size_t)w * dim + col] = run;
                run += __popc(bitmap[(size_t)w * dim + col]);
            }
            colcount[col] = run;
        }
    }
}

// host side: per-segment group ids / last flags and the slow-group table, built on first use
int csc_tile_prepare(ofb_plan *p) {
    if (p->d_seg_group) return OFB_OK;
    OFB_TRY
    const int64_t G = p->n_groups;
    const size_t n_segs = p->h_seg_begin.empty() ? 0 : p->h_seg_begin[G];
    std::vector<uint32_t> seg_group(std::max<size_t>(n_segs, 1), 0);
    for (int64_t g = 0; g < G; ++g)
        for (uint32_t s = p->h_seg_begin[g]; s < p->h_seg_begin[g + 1]; ++s) seg_group[s] = (uint32_t)g;
    OFB_CUDA(cudaMalloc(&p->d_seg_group, seg_group.size() * sizeof(uint32_t)));
    OFB_CUDA(cudaMemcpy(p->d_seg_group, seg_group.data(), seg_group.size() * sizeof(uint32_t),
                        cudaMemcpyHostToDevice));
    OFB_CUDA(cudaMalloc(&p->d_group_slow, std::max<size_t>(p->h_group_slow.size(), 1)));
    if (!p->h_group_slow.empty())
        OFB_CUDA(cudaMemcpy(p->d_group_slow, p->h_group_slow.data(), p->h_group_slow.size(),
                            cudaMemcpyHostToDevice));
    return OFB_OK;
    OFB_CATCH((void)0)
}

bool csc_tile_ok(const ofb_plan *p) {
    return !p->projected && p->n_qubits >= TTILE_BITS && p->n_qubits <= 30 && p->n_groups > 0 &&
           p->d_segs != nullptr;
}

int csc_tile_count(ofb_plan *p, uint32_t words, uint32_t *bitmap, uint32_t *prefix, int64_t *counts,
                   cudaStream_t s) {
    int rc = csc_tile_prepare(p);
    if (rc) return rc;
    const uint64_t dim = 1ull << p->n_qubits;
    OFB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)words * dim * sizeof(uint32_t), s));
    const unsigned grid = (unsigned)(dim / TTILE);
#define OFB_TILE(II)                                                                              \
    k_csc_tile<false, II, int32_t><<<grid, TB, 0, s>>>(                                           \
        p->d_chunks, (uint32_t)p->n_chunks, p->d_segs, p->d_seg_group, p->d_zc, p->d_zc_im, p->d_sib, \
        p->d_group_slow, p->d_rows, p->d_groups, p->n_qubits, dim, words, bitmap, prefix, counts,  \
        nullptr, (int32_t *)nullptr, nullptr)
    if (p->has_imag) OFB_TILE(true); else OFB_TILE(false);
#undef OFB_TILE
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int csc_tile_fill(ofb_plan *p, uint32_t words, uint32_t *bitmap, uint32_t *prefix, const int64_t *colptr,
                  void *d_indices, void *d_data, int index_bits, cudaStream_t s) {
    const uint64_t dim = 1ull << p->n_qubits;
    const unsigned grid = (unsigned)(dim / TTILE);
#define OFB_TILE(II, TT)                                                                          \
    k_csc_tile<true, II, TT><<<grid, TB, 0, s>>>(                                                 \
        p->d_chunks, (uint32_t)p->n_chunks, p->d_segs, p->d_seg_group, p->d_zc, p->d_zc_im, p->d_sib, \
        p->d_group_slow, p->d_rows, p->d_groups, p->n_qubits, dim, words, bitmap, prefix, nullptr, \
        colptr, (TT *)d_indices, (double2 *)d_data)
    if (index_bits == 32) {
        if (p->has_imag) OFB_TILE(true, int32_t); else OFB_TILE(false, int32_t);
    } else {
        if (p->has_imag) OFB_TILE(true, int64_t); else OFB_TILE(false, int64_t);
    }
#undef OFB_TILE
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

}  // namespace ofb
