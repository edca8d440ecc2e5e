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
#ifndef OFB_CSC_QW
#define OFB_CSC_QW 4
#endif
constexpr int QW = OFB_CSC_QW;  // columns whose position lookups are in flight together
static_assert(CLASS_BITS == 4 && CLASS_SHIFT == 6, "csc_tile.cu assumes the 64 x 16 geometry");

struct CscStage {
    Segment segs[CHUNK_MAX_SEGS];
    TermZC re[CHUNK_MAX_TERMS + 8];
    TermZC im[CHUNK_MAX_TERMS + 8];
};

struct __align__(16) CscSlice {
    uint32_t ch0, ch1, seg0, seg1;  // chunks and segments of the slice (whole groups)
    uint32_t term0, term1, pad0, pad1;  // absolute term range
    Chunk first;                        // chunks[ch0]: saves the CTA one dependent load
};

__device__ __forceinline__ void csc_cp16(void *smem, const void *gmem) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}

// Stages the part of chunk `ck` that belongs to the slice: segments [seg0, seg1) and terms
// [term0, term1) (absolute indices), at their usual positions inside the buffer.  A slice of 2-3
// groups does not pay for the 16 segments / 256 terms of a whole chunk.
__device__ __forceinline__ void csc_stage(CscStage &b, const Chunk ck, const Segment *__restrict__ segs,
                                          const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im,
                                          bool has_imag, uint32_t seg0, uint32_t seg1, uint32_t term0,
                                          uint32_t term1) {
    const uint32_t sa = max(ck.seg_begin, seg0) - ck.seg_begin;
    const uint32_t sb = min(ck.seg_begin + ck.seg_count, seg1) - ck.seg_begin;
    const char *gs = (const char *)(segs + ck.seg_begin);
    char *ss = (char *)b.segs;
    for (uint32_t u = 2 * sa + threadIdx.x; u < 2 * sb; u += TB) csc_cp16(ss + u * 16, gs + u * 16);
    const uint32_t ta = max(ck.term_begin, term0) - ck.term_begin;
    const uint32_t tb = min(ck.term_begin + ck.term_count, term1) - ck.term_begin;
    for (uint32_t u = ta + threadIdx.x; u < tb; u += TB) {
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

// blockIdx.x = tile * n_slices + slice: a tile of 1 024 columns and a slice of the group list
// (whole groups in x order; see csc_slice_count).  The slices of a tile are neighbours in launch
// order, so the tiles being written at any time are few.
// slice = (first chunk, end chunk, first segment, end segment, first term, end term).
template <bool FILL, bool HAS_IMAG, typename I>
__global__ void __launch_bounds__(TB, 6)
k_csc_tile(const Chunk *__restrict__ chunks, const CscSlice *__restrict__ slices, uint32_t n_slices,
           const Segment *__restrict__ segs,
           const uint32_t *__restrict__ seg_group, const TermZC *__restrict__ zc_re,
           const TermZC *__restrict__ zc_im, const uint32_t *__restrict__ sib,
           const uint8_t *__restrict__ group_slow, const ScatterRow *__restrict__ rows,
           const GroupDesc *__restrict__ groups, int n, uint64_t dim, uint32_t words,
           uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ prefix,
           const I *__restrict__ colptr, I *__restrict__ indices, double2 *__restrict__ data) {
    __shared__ CscStage sbuf[2];
    __shared__ uint32_t s_sib[2][CHUNK_MAX_SEGS][32];
    __shared__ uint32_t s_group[2][CHUNK_MAX_SEGS];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t tile = blockIdx.x / n_slices;
    const uint32_t c0 = tile * (uint32_t)TTILE + tid;  // column of r = 0

    double vr[TR], vi[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) vr[r] = vi[r] = 0.0;
#ifndef OFB_CSC_PLAIN_STORES
    uint64_t l2_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_keep));
#endif

    const CscSlice slice = slices[blockIdx.x - tile * n_slices];
    const uint32_t ch0 = slice.ch0, ch1 = slice.ch1, seg0 = slice.seg0, seg1 = slice.seg1;
    const uint32_t term0 = slice.term0, term1 = slice.term1;

    auto stage_aux = [&](int buf, const Chunk ck) {  // group ids and sibling rows of the slice's segments
        const uint32_t qa = max(ck.seg_begin, seg0) - ck.seg_begin;
        const uint32_t qb = min(ck.seg_begin + ck.seg_count, seg1) - ck.seg_begin;
        for (uint32_t q = qa; q < qb; ++q) {
            const uint32_t g = seg_group[ck.seg_begin + q];
            if (tid == 0) s_group[buf][q] = g;
            if (tid < 32) s_sib[buf][q][tid] = tid < (uint32_t)n ? sib[(size_t)g * n + tid] : 0u;
        }
    };

    Chunk cur = slice.first;
    Chunk nxt = ch0 + 1 < ch1 ? chunks[ch0 + 1] : cur;
    csc_stage(sbuf[0], cur, segs, zc_re, zc_im, HAS_IMAG, seg0, seg1, term0, term1);
    stage_aux(0, cur);
    for (uint32_t cc = ch0; cc < ch1; ++cc) {
        const uint32_t c = cc - ch0;  // buffer parity
        Chunk after = nxt;
        if (cc + 1 < ch1) {
            csc_stage(sbuf[(c + 1) & 1], nxt, segs, zc_re, zc_im, HAS_IMAG, seg0, seg1, term0, term1);
            stage_aux((c + 1) & 1, nxt);
            if (cc + 2 < ch1) after = chunks[cc + 2];
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const CscStage &b = sbuf[c & 1];
        for (uint32_t q = 0; q < cur.seg_count; ++q) {
            if (cur.seg_begin + q < seg0 || cur.seg_begin + q >= seg1) continue;  // another slice's groups
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
            // Four columns at a time: the position lookups (bitmap word, prefix count, column start)
            // of all four are issued before the first store, so their latencies overlap -- with one
            // column at a time the pass was bound by 12 warps per SM each waiting on one load chain
            // (ncu: 33 warp-cycles of long-scoreboard stall per issued instruction).
#pragma unroll
            for (int r0 = 0; r0 < TR; r0 += QW) {
                bool nz[QW];
                int32_t rank[QW];
                uint32_t bm[QW], pf[QW];
                I cp[QW];
#pragma unroll
                for (int q = 0; q < QW; ++q) {
                    const int r = r0 + q;
                    nz[q] = !(vr[r] == 0.0 && vi[r] == 0.0);
                    rank[q] = (int32_t)base;
#pragma unroll
                    for (int k = 0; k < CLASS_BITS; ++k) rank[q] += ((r >> k) & 1) ? delta[k] : 0;
                    const uint64_t col = (uint64_t)(c0 + r * TB);
                    const size_t at = (size_t)((uint32_t)rank[q] >> 5) * dim + col;
                    if (!FILL) {
                        if (nz[q]) atomicOr(&bitmap[at], 1u << (rank[q] & 31));  // other slices set other bits
                    } else {
                        bm[q] = nz[q] ? bitmap[at] : 0u;
                        pf[q] = nz[q] ? prefix[at] : 0u;
                        cp[q] = nz[q] ? colptr[col] : 0;
                    }
                }
#pragma unroll
                for (int q = 0; q < QW; ++q) {
                    const int r = r0 + q;
                    if (FILL && nz[q]) {
                        const uint64_t col = (uint64_t)(c0 + r * TB);
                        const int64_t pos = (int64_t)cp[q] + pf[q] + __popc(bm[q] & ((1u << (rank[q] & 31)) - 1u));
#ifndef OFB_CSC_PLAIN_STORES
                        // partially written sectors should outlive the lookup traffic in L2: stores carry an
                        // evict_last policy (measured: fill 44.4 -> 38.2 ms at S = 32, 34.4 -> 30.6 at S = 64,
                        // 15.15 -> 14.97 ms at the default S; never slower)
                        const I row = (I)(col ^ (uint64_t)x);
                        if (sizeof(I) == 4)
                            asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(indices + pos), "r"((uint32_t)row), "l"(l2_keep) : "memory");
                        else
                            asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(indices + pos), "l"((uint64_t)row), "l"(l2_keep) : "memory");
                        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(data + pos), "d"(vr[r]), "d"(vi[r]), "l"(l2_keep) : "memory");
#else
                        indices[pos] = (I)(col ^ (uint64_t)x);
                        data[pos] = make_double2(vr[r], vi[r]);
#endif
                    }
                    vr[r] = 0.0;
                    vi[r] = 0.0;
                }
            }
        }
        __syncthreads();
        cur = nxt;
        nxt = after;
    }
}

// per-column prefix counts of the rank bitmap (after ALL slices of the count pass): prefix[w][c] =
// non-zeros of column c in words < w, colcount[c] = their total.  One thread per column, coalesced.
__global__ void __launch_bounds__(256)
k_csc_prefix(uint64_t dim, uint32_t words, const uint32_t *__restrict__ bitmap, uint32_t *__restrict__ prefix,
             int64_t *__restrict__ colcount) {
    const uint64_t col = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    uint32_t run = 0;
    for (uint32_t w = 0; w < words; ++w) {
        prefix[(size_t)w * dim + col] = run;
        run += __popc(bitmap[(size_t)w * dim + col]);
    }
    colcount[col] = run;
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

// Slices of the group list (whole groups in x order, balanced by terms + a per-group constant for
// the rank and store work), one CTA per (tile, slice).  Two reasons to cut:
// (1) a small matrix has too few column tiles to fill 148 SMs x 6 CTAs;
// (2) the fill pass appends 16 + 4 bytes per entry to 1 024 columns whose output ranges lie
//     kilobytes apart, so every 32-byte sector is written in 2 (data) or 8 (indices) separate
//     stores and is complete only if it stays in L2 until the last of them.  With one CTA per
//     tile the 888 resident CTAs have their whole output open at once (14.9 GB at config 2): ncu
//     showed 39 GB written + 35 GB read for those 14.9 GB -- sectors evicted half written and
//     read back.  With S slices per tile only 888 / S tiles are open; measured on B200 (with a
//     CTA staging only its own slice of a chunk) the pass is fastest when the open output is
//     ~10 MB, down to one group per slice (config 2: S = 1 024, 54.9 -> 15.3 ms; 24-qubit
//     Hubbard: S = 49 = G, 25.9 -> 6.0 ms).  Visiting the groups in ROW order instead (blocks of
//     equal high x bits sorted by key ^ c_hi per tile, so that each column advances monotonically)
//     was built and measured: no gain at S = 1 (51.5 ms: an index sector still waits for 8
//     entries of its column while the chip writes 165 MB) and slower than x-order slices at equal
//     S, so it was dropped (profiles/README.md).
// OFB_CSC_COUNT_SLICES / OFB_CSC_FILL_SLICES / OFB_CSC_SLICES override the choice.
static uint32_t csc_slice_count(const ofb_plan *p, bool fill, double bytes_per_tile) {
    const char *env = getenv(fill ? "OFB_CSC_FILL_SLICES" : "OFB_CSC_COUNT_SLICES");
    if (!env) env = getenv("OFB_CSC_SLICES");
    const uint32_t G = (uint32_t)p->n_groups;
    if (env) return std::min<uint32_t>(std::max(atoi(env), 1), G);
    const double tiles = (double)((1ull << p->n_qubits) / TTILE);
    const double resident = 148.0 * 6.0;
    double s = std::ceil(4.0 * resident / tiles);  // (1) at least ~4 waves of CTAs
    if (fill && bytes_per_tile > 0.0) s = std::max(s, std::ceil(resident * bytes_per_tile / 10e6));  // (2)
    s = std::min(s, (double)G);  // at least one group per slice
    return (uint32_t)std::min<double>(std::max(s, 1.0), 4096.0);
}

static int csc_tile_slices(ofb_plan *p, uint32_t want, void **d_out, uint32_t *n_out) {
    OFB_TRY
    const uint32_t G = (uint32_t)p->n_groups;
    want = std::min(std::max(want, 1u), G);
    if (*d_out && *n_out == want) return OFB_OK;
    // weight of a group: its terms + 16 (rank computation and up to 16 stores per thread)
    std::vector<double> acc(G + 1, 0.0);
    for (uint32_t g = 0; g < G; ++g)
        acc[g + 1] = acc[g] + (double)(p->h_group_begin[g + 1] - p->h_group_begin[g]) + 16.0;
    std::vector<CscSlice> slices;
    uint32_t g0 = 0;
    for (uint32_t k = 1; k <= want; ++k) {
        // slice k ends at the first group boundary at or after its share of the weight, but
        // leaves at least one group for every slice still to come
        uint32_t g1 = G;
        if (k < want) {
            const double target = acc[G] * k / want;
            g1 = (uint32_t)(std::lower_bound(acc.begin(), acc.end(), target) - acc.begin());
            g1 = std::min(std::max(g1, g0 + 1), G - (want - k));
        }
        const uint32_t seg0 = p->h_seg_begin[g0], seg1 = p->h_seg_begin[g1];
        slices.push_back(CscSlice{p->h_chunk_of_seg[seg0], p->h_chunk_of_seg[seg1 - 1] + 1, seg0, seg1,
                                  p->h_group_begin[g0], p->h_group_begin[g1], 0u, 0u,
                                  p->h_chunks[p->h_chunk_of_seg[seg0]]});
        g0 = g1;
    }
    cudaFree(*d_out);
    *d_out = nullptr;
    OFB_CUDA(cudaMalloc(d_out, slices.size() * sizeof(CscSlice)));
    OFB_CUDA(cudaMemcpy(*d_out, slices.data(), slices.size() * sizeof(CscSlice), cudaMemcpyHostToDevice));
    *n_out = (uint32_t)slices.size();
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
    if (!rc) rc = csc_tile_slices(p, csc_slice_count(p, false, 0.0), &p->d_count_slices, &p->n_count_slices);
    if (rc) return rc;
    const uint64_t dim = 1ull << p->n_qubits;
    OFB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)words * dim * sizeof(uint32_t), s));
    const uint32_t n_slices = p->n_count_slices;
    const unsigned grid = (unsigned)(dim / TTILE) * n_slices;
#define OFB_TILE(II)                                                                              \
    k_csc_tile<false, II, int32_t><<<grid, TB, 0, s>>>(                                           \
        p->d_chunks, (const CscSlice *)p->d_count_slices, n_slices, p->d_segs, p->d_seg_group, p->d_zc, p->d_zc_im, p->d_sib,       \
        p->d_group_slow, p->d_rows, p->d_groups, p->n_qubits, dim, words, bitmap, prefix,         \
        (const int32_t *)nullptr, (int32_t *)nullptr, nullptr)
    if (p->has_imag) OFB_TILE(true); else OFB_TILE(false);
#undef OFB_TILE
    OFB_LAUNCH_CHECK();
    k_csc_prefix<<<(unsigned)((dim + 255) / 256), 256, 0, s>>>(dim, words, bitmap, prefix, counts);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int csc_tile_fill(ofb_plan *p, uint32_t words, uint32_t *bitmap, uint32_t *prefix, const void *d_indptr,
                  void *d_indices, void *d_data, int index_bits, cudaStream_t s) {
    const uint64_t dim = 1ull << p->n_qubits;
    const double bytes_per_tile = (double)p->csc_nnz * (16.0 + index_bits / 8) / (double)(dim / TTILE);
    int rc = csc_tile_slices(p, csc_slice_count(p, true, bytes_per_tile), &p->d_fill_slices, &p->n_fill_slices);
    if (rc) return rc;
    const uint32_t n_slices = p->n_fill_slices;
    const unsigned grid = (unsigned)(dim / TTILE) * n_slices;
#define OFB_TILE(II, TT)                                                                          \
    k_csc_tile<true, II, TT><<<grid, TB, 0, s>>>(                                                 \
        p->d_chunks, (const CscSlice *)p->d_fill_slices, n_slices, p->d_segs, p->d_seg_group, p->d_zc, p->d_zc_im, p->d_sib,       \
        p->d_group_slow, p->d_rows, p->d_groups, p->n_qubits, dim, words, bitmap, prefix,         \
        (const TT *)d_indptr, (TT *)d_indices, (double2 *)d_data)
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
