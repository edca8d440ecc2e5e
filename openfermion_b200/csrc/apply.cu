// K1 matrix-free H|psi>, K2 fused expectation, K3 diagonal.
//
// Gather form:
//     out[j] = sum_g S_g(j) * in[j ^ x_g],   S_g(j) = sum_{t in g} c_t (-1)^popc(j & z_t)
//
// Main kernel (k_apply_tile): a CTA of B = 2^BLOCK_BITS threads (64) owns a tile of RT * B
// amplitudes (RT = 2^CLASS_BITS = 16 per thread); thread `tid` owns j_r = tile + r*B + tid,
// r = 0..RT-1 (stride B, so every 128-bit load/store instruction of a warp covers 512 contiguous
// bytes and XOR with x only permutes lanes inside aligned blocks: loads stay sector-coalesced
// for any mask).  The RT amplitudes of a thread differ only in index bits s..s+CLASS_BITS-1
// (s = log2 B = CLASS_SHIFT), so the sign of term t on amplitude r is
//     sigma_t(j_0) * (-1)^popc(r & w_t),      w_t = (z_t >> s) & CLASS_MASK   (the term's "class").
// The plan stores each x-group's terms sorted by class (segments of <= 255 terms and <= 8 class
// runs, run lengths and class ids packed in the descriptor), so per class run a thread
// accumulates
//     A = sum_{t in run} sigma_t(j_0) c_t     -- LDS.128 + LOP3 + POPC + IMAD + DADD per term,
// software pipelined over pairs of terms and across runs -- and applies A to its RT amplitudes
// with 2*RT DFMA whose +- pattern is a compile-time constant per class (a warp-uniform switch).
// Groups whose phased coefficients are all real (or all imaginary) cost 2 DFMA per amplitude;
// complex groups run the imaginary table as a second pass over the same loaded amplitudes.
// Groups are executed in ascending x, so consecutive groups re-read the same input tile through
// L1/L2; term tables are staged to shared memory chunk by chunk with cp.async (double buffered).
// Per (segment, amplitude) this issues 40 % fewer instructions than the 8-amplitude version
// (the per-term work is shared by twice as many amplitudes) at 12 instead of 16 warps per SM.
#include <algorithm>
#include <cstdlib>

#include "ofb_internal.cuh"

namespace ofb {

constexpr int BLOCK = 1 << OFB_APPLY_BLOCK_BITS;
constexpr int RT = 1 << CLASS_BITS;  // amplitudes per thread (8, or 16 with OFB_APPLY_CLASS_BITS=4)
constexpr int TILE = BLOCK * RT;
constexpr int TILE_BITS = OFB_APPLY_BLOCK_BITS + CLASS_BITS;
constexpr int HOST_SLABS = 8;  // ofb_apply_host: output slabs copied back while later tiles run
static_assert(BLOCK == (1 << CLASS_SHIFT), "class bits must sit right above the thread index");

__device__ __forceinline__ double signed_coeff(double c, uint32_t parity) {
    return __hiloint2double(__double2hiint(c) ^ (int)(parity << 31), __double2loint(c));
}

template <bool Z32>
__device__ __forceinline__ uint32_t sign_parity(uint64_t z, uint32_t j_lo, uint32_t base_hi) {
    // parity of popc((index_base | j) & z); for n <= 32 everything lives in the low word
    if (Z32) return (uint32_t)__popc(j_lo & (uint32_t)z);
    return (uint32_t)(__popc(j_lo & (uint32_t)z) + __popc(base_hi & (uint32_t)(z >> 32)));
}

// sigma-signed sum of `cnt` consecutive table entries (one class run).  The sign is folded
// into the high word with one add: hi + (parity << 31) flips bit 31 (mod 2^32).  Two entries
// per iteration on independent accumulators, software pipelined: the caller passes the first
// pair already loaded (`e0`, `e1`, issued before the previous run's FMAs) and every iteration
// issues the loads of the next pair before consuming the current one, so the shared-memory
// latency overlaps the popc/add chains even at three warps per scheduler.  Reads may run up to
// two entries past the run (the staging buffers are padded); those values are never used.
template <bool Z32>
__device__ __forceinline__ double signed_entry(const TermZC &e, uint32_t j_lo, uint32_t base_hi) {
    return __hiloint2double(__double2hiint(e.c) + (int)(sign_parity<Z32>(e.z, j_lo, base_hi) << 31),
                            __double2loint(e.c));
}

template <bool Z32>
__device__ __forceinline__ double run_sum(const TermZC *tab, uint32_t cnt, TermZC e0, TermZC e1,
                                          uint32_t j_lo, uint32_t base_hi) {
    double a0 = 0.0, a1 = 0.0;
    const TermZC *p = tab + 2;
    const TermZC *end2 = tab + (cnt & ~1u);
    if (cnt >= 2u) {
#pragma unroll 1
        while (p != end2) {
            const TermZC n0 = p[0], n1 = p[1];
            p += 2;
            a0 += signed_entry<Z32>(e0, j_lo, base_hi);
            a1 += signed_entry<Z32>(e1, j_lo, base_hi);
            e0 = n0;
            e1 = n1;
        }
        a0 += signed_entry<Z32>(e0, j_lo, base_hi);
        a1 += signed_entry<Z32>(e1, j_lo, base_hi);
        if (cnt & 1u) a0 += signed_entry<Z32>(end2[0], j_lo, base_hi);
    } else {  // cnt == 1: the preloaded first entry is the whole run
        a0 = signed_entry<Z32>(e0, j_lo, base_hi);
    }
    return a0 + a1;
}

// out_r += (+-a) * v_r (real pass) or (+- i a) * v_r (imaginary pass) with the static sign
// (-1)^popc(r & W) of class W on amplitude r folded into the FMA operand modifiers.
template <int W, bool IMAG>
__device__ __forceinline__ void apply_class(double a, const double2 (&v)[RT], double (&ar)[RT],
                                            double (&ai)[RT]) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const bool neg = (__builtin_popcount(r & W) & 1) != 0;
        const double s = neg ? -a : a;
        if (!IMAG) {
            ar[r] = fma(s, v[r].x, ar[r]);
            ai[r] = fma(s, v[r].y, ai[r]);
        } else {  // (i s)(vx + i vy) = -s vy + i s vx
            ar[r] = fma(-s, v[r].y, ar[r]);
            ai[r] = fma(s, v[r].x, ai[r]);
        }
    }
}

template <bool IMAG>
__device__ __forceinline__ void apply_class_dyn(uint32_t w, double a, const double2 (&v)[RT],
                                                double (&ar)[RT], double (&ai)[RT]) {
    switch (w) {
        case 0: apply_class<0, IMAG>(a, v, ar, ai); break;
        case 1: apply_class<1, IMAG>(a, v, ar, ai); break;
        case 2: apply_class<2, IMAG>(a, v, ar, ai); break;
        case 3: apply_class<3, IMAG>(a, v, ar, ai); break;
        case 4: apply_class<4, IMAG>(a, v, ar, ai); break;
        case 5: apply_class<5, IMAG>(a, v, ar, ai); break;
        case 6: apply_class<6, IMAG>(a, v, ar, ai); break;
#if OFB_APPLY_CLASS_BITS == 4
        case 7: apply_class<7, IMAG>(a, v, ar, ai); break;
        case 8: apply_class<8, IMAG>(a, v, ar, ai); break;
        case 9: apply_class<9, IMAG>(a, v, ar, ai); break;
        case 10: apply_class<10, IMAG>(a, v, ar, ai); break;
        case 11: apply_class<11, IMAG>(a, v, ar, ai); break;
        case 12: apply_class<12, IMAG>(a, v, ar, ai); break;
        case 13: apply_class<13, IMAG>(a, v, ar, ai); break;
        case 14: apply_class<14, IMAG>(a, v, ar, ai); break;
        default: apply_class<15, IMAG>(a, v, ar, ai); break;
#else
        default: apply_class<7, IMAG>(a, v, ar, ai); break;
#endif
    }
}

// all class runs of one segment from one coefficient table
template <bool Z32, bool IMAG>
__device__ __forceinline__ void segment_pass(const TermZC *tab, uint32_t runs, uint64_t cnts,
                                             uint32_t ids, uint32_t j_lo, uint32_t base_hi,
                                             const double2 (&v)[RT], double (&ar)[RT],
                                             double (&ai)[RT]) {
    TermZC e0 = tab[0], e1 = tab[1];
    if (runs == 1) {  // the common case: no run bookkeeping at all
        const double a = run_sum<Z32>(tab, (uint32_t)cnts, e0, e1, j_lo, base_hi);
        apply_class_dyn<IMAG>(ids, a, v, ar, ai);
        return;
    }
#pragma unroll 1
    for (uint32_t k = 0; k < runs; ++k) {
        const uint32_t cnt = (uint32_t)cnts & 0xffu;
        const double a = run_sum<Z32>(tab, cnt, e0, e1, j_lo, base_hi);
        tab += cnt;
        // first pair of the next run: in flight while this run's FMAs execute
        e0 = tab[0];
        e1 = tab[1];
        apply_class_dyn<IMAG>(ids & 0xfu, a, v, ar, ai);
        cnts >>= 8;
        ids >>= 4;
    }
}

// The RT input amplitudes of a thread for one segment.  (A variant that computes one address
// and dispatches on the class bits of x to issue the loads with compile-time offsets was
// measured 8 % slower: the branch keeps the loads from being issued early.)
__device__ __forceinline__ void load_tile(const double2 *__restrict__ vin, uint32_t jx, double2 (&v)[RT]) {
#pragma unroll
    for (int r = 0; r < RT; ++r) v[r] = vin[(size_t)(jx ^ (uint32_t)(r << CLASS_SHIFT))];
}

// Same with a support hint (ofb_plan_set_support): amplitudes whose coefficient S_g(j) is
// structurally zero are not fetched; they enter the sums as exact zeros.
__device__ __forceinline__ void load_tile_needed(const double2 *__restrict__ vin, uint32_t jx,
                                                 uint32_t need, double2 (&v)[RT]) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        v[r] = make_double2(0.0, 0.0);
        if ((need >> r) & 1u) v[r] = vin[(size_t)(jx ^ (uint32_t)(r << CLASS_SHIFT))];
    }
}

// Shared-memory stage of one chunk: segment descriptors and both coefficient tables.
struct StageBuf {
    Segment segs[CHUNK_MAX_SEGS];
    TermZC re[CHUNK_MAX_TERMS + 8];  // +8: the prefetching reader runs one entry ahead
    TermZC im[CHUNK_MAX_TERMS + 8];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <bool COND>
__device__ __forceinline__ void stage_chunk(StageBuf &b, SegCond *sc, const Chunk ck,
                                            const Segment *__restrict__ segs,
                                            const SegCond *__restrict__ conds,
                                            const TermZC *__restrict__ zc_re,
                                            const TermZC *__restrict__ zc_im, int has_imag) {
    if (COND)
        for (uint32_t u = threadIdx.x; u < ck.seg_count; u += BLOCK) cp_async16(sc + u, conds + ck.seg_begin + u);
    const char *gs = (const char *)(segs + ck.seg_begin);
    char *ss = (char *)b.segs;
    for (uint32_t u = threadIdx.x; u < ck.seg_count * 2; u += BLOCK) cp_async16(ss + u * 16, gs + u * 16);
    for (uint32_t u = threadIdx.x; u < ck.term_count; u += BLOCK) {
        cp_async16(&b.re[u], zc_re + ck.term_begin + u);
        if (has_imag) cp_async16(&b.im[u], zc_im + ck.term_begin + u);
    }
    cp_async_commit();
}

// MODE 0: write H in; 1: fused expectation conj(in) . (H in), no output vector (K2);
// 2: write H in and the per-tile partial sums of conj(vdot) . (H in) -- the Lanczos alpha for free.
template <bool Z32, int MODE, bool HAS_IMAG, bool COND>
__global__ void __launch_bounds__(BLOCK, OFB_APPLY_MIN_CTAS)
k_apply_tile(const double2 *__restrict__ vin, double2 *__restrict__ vout,
         const Chunk *__restrict__ chunks, uint32_t c0, uint32_t c1,
         const Segment *__restrict__ segs, uint32_t s0, uint32_t s1,
         const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im, int has_imag,
         uint32_t xmask_local, uint64_t index_base, int accumulate, double2 *__restrict__ partials,
         const SegCond *__restrict__ conds, uint32_t tile0, const double2 *__restrict__ vdot) {
    constexpr bool EXPECT = MODE == 1;
    __shared__ StageBuf sbuf[2];
    __shared__ SegCond scond[COND ? 2 : 1][COND ? CHUNK_MAX_SEGS : 1];
    const uint32_t tile = blockIdx.x + tile0;  // a launch may cover a slab of the tiles
    const uint32_t j0 = tile * (uint32_t)TILE + threadIdx.x;  // local index of r = 0
    const uint32_t j_lo = (uint32_t)index_base | j0;
    const uint32_t base_hi = (uint32_t)(index_base >> 32);

    double ar[RT], ai[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        ar[r] = 0.0;
        ai[r] = 0.0;
    }
    if (!EXPECT && accumulate) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            double2 o = vout[(size_t)(j0 + r * BLOCK)];
            ar[r] = o.x;
            ai[r] = o.y;
        }
    }

    Chunk cur = chunks[c0];
    Chunk nxt = (c0 + 1 < c1) ? chunks[c0 + 1] : cur;
    stage_chunk<COND>(sbuf[0], scond[0], cur, segs, conds, zc_re, zc_im, HAS_IMAG ? 1 : 0);
    for (uint32_t c = c0; c < c1; ++c) {
        Chunk after = nxt;
        if (c + 1 < c1) {
            stage_chunk<COND>(sbuf[(c + 1 - c0) & 1], scond[COND ? ((c + 1 - c0) & 1) : 0], nxt, segs, conds,
                              zc_re, zc_im, HAS_IMAG ? 1 : 0);
            if (c + 2 < c1) after = chunks[c + 2];
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const StageBuf &b = sbuf[(c - c0) & 1];
        const uint32_t lo = max(cur.seg_begin, s0), hi = min(cur.seg_begin + cur.seg_count, s1);
        for (uint32_t s = lo; s < hi; ++s) {
            const Segment &sd = b.segs[s - cur.seg_begin];
            const uint32_t kind_runs = sd.kind_runs;
            const uint32_t ids = sd.ids;
            const uint64_t cnts = sd.cnts;
            const uint32_t tl = sd.rel_begin;
            const uint32_t jx = j0 ^ (sd.x_lo & xmask_local);
            const uint32_t runs = kind_runs & 0xfu;
            double2 v[RT];
            if (COND) {
                // which of this thread's amplitudes can have a non-zero coefficient in this group
                const SegCond cd = scond[(c - c0) & 1][s - cur.seg_begin];
                const uint32_t q1 = ((uint32_t)__popc(j_lo & cd.m1) + cd.v) & 1u;
                const uint32_t q2 = ((uint32_t)__popc(j_lo & cd.m2) + (cd.v >> 1)) & 1u;
                const uint32_t need = (cd.pat ^ (q1 - 1u)) & ((cd.pat >> 16) ^ (q2 - 1u)) & ((1u << RT) - 1u);
                if (need == 0u) continue;
                // conditions that do not involve the class bits leave a thread all or nothing:
                // the common case takes the plain loads (immediate offsets, no predicates)
                if (need == ((1u << RT) - 1u))
                    load_tile(vin, jx, v);
                else
                    load_tile_needed(vin, jx, need, v);
            } else {
                load_tile(vin, jx, v);
            }
            if (!HAS_IMAG) {
                segment_pass<Z32, false>(b.re + tl, runs, cnts, ids, j_lo, base_hi, v, ar, ai);
            } else {
                const uint32_t kind = kind_runs >> 30;
                if (kind != KIND_IMAG)
                    segment_pass<Z32, false>(b.re + tl, runs, cnts, ids, j_lo, base_hi, v, ar, ai);
                if (kind != KIND_REAL)
                    segment_pass<Z32, true>(b.im + tl, runs, cnts, ids, j_lo, base_hi, v, ar, ai);
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
        cur = nxt;
        nxt = after;
    }

    if (MODE != 1) {
#pragma unroll
        for (int r = 0; r < RT; ++r) vout[(size_t)(j0 + r * BLOCK)] = make_double2(ar[r], ai[r]);
    }
    if (MODE != 0) {
        // conj(w[j]) * (H in)[j], w = in (expectation) or vdot, block-reduced in a fixed order
        const double2 *__restrict__ wv = MODE == 1 ? vin : vdot;
        double er = 0.0, ei = 0.0;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            double2 w = wv[(size_t)(j0 + r * BLOCK)];
            er += w.x * ar[r] + w.y * ai[r];
            ei += w.x * ai[r] - w.y * ar[r];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            er += __shfl_xor_sync(0xffffffffu, er, off);
            ei += __shfl_xor_sync(0xffffffffu, ei, off);
        }
        __shared__ double sm[2 * (BLOCK / 32)];
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            sm[2 * warp] = er;
            sm[2 * warp + 1] = ei;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tr = 0.0, ti = 0.0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; ++w) {
                tr += sm[2 * w];
                ti += sm[2 * w + 1];
            }
            partials[tile] = make_double2(tr, ti);
        }
    }
}

// K1b: H applied to several columns at once (scipy's _matmat behind the block Davidson,
// linalg/davidson.py:249-256).  The sign sums A of a class run depend on the row index only, not
// on the vector, so they are computed ONCE per thread for a batch of segments, parked in shared
// memory (a thread reads back only its own slots, so no barrier is needed), and then applied to
// every column: per (segment, column) a thread issues its 16 gathers and the multiply-adds, the
// per-term work (LDS + LOP3 + POPC + IMAD + DADD) and the support test are shared by the columns.
// The accumulators of one column live in registers only while that column is being processed;
// between batches they rest in the output vector itself (one extra 16-amplitude read-modify-write
// per column per batch of ~14 segments: < 10 % of the gather traffic).
constexpr int MB_RUNS = 24;  // class runs per batch

template <bool Z32>
__device__ __forceinline__ void store_run_sums(const TermZC *tab, uint32_t runs, uint64_t cnts,
                                               uint32_t j_lo, uint32_t base_hi, double *slot) {
#pragma unroll 1
    for (uint32_t k = 0; k < runs; ++k) {
        const uint32_t cnt = (uint32_t)cnts & 0xffu;
        slot[k * BLOCK] = run_sum<Z32>(tab, cnt, tab[0], tab[1], j_lo, base_hi);
        tab += cnt;
        cnts >>= 8;
    }
}

template <bool Z32, bool HAS_IMAG, bool COND>
__global__ void __launch_bounds__(BLOCK, 4)
k_apply_multi(const double2 *__restrict__ vin, double2 *__restrict__ vout, int ncols, int64_t ld,
              const Chunk *__restrict__ chunks, uint32_t c0, uint32_t c1,
              const Segment *__restrict__ segs, uint32_t s0, uint32_t s1,
              const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im,
              uint32_t xmask_local, uint64_t index_base, const SegCond *__restrict__ conds) {
    __shared__ StageBuf sbuf[2];
    __shared__ SegCond scond[COND ? 2 : 1][COND ? CHUNK_MAX_SEGS : 1];
    __shared__ double a_re[MB_RUNS][BLOCK];
    __shared__ double a_im[HAS_IMAG ? MB_RUNS : 1][BLOCK];
    __shared__ uint16_t s_need[COND ? CHUNK_MAX_SEGS : 1][BLOCK];
    const uint32_t tid = threadIdx.x;
    const uint32_t j0 = blockIdx.x * (uint32_t)TILE + tid;
    const uint32_t j_lo = (uint32_t)index_base | j0;
    const uint32_t base_hi = (uint32_t)(index_base >> 32);
    constexpr uint32_t ALL = (1u << RT) - 1u;
    bool first_batch = true;

    Chunk cur = chunks[c0];
    Chunk nxt = (c0 + 1 < c1) ? chunks[c0 + 1] : cur;
    stage_chunk<COND>(sbuf[0], scond[0], cur, segs, conds, zc_re, zc_im, HAS_IMAG ? 1 : 0);
    for (uint32_t c = c0; c < c1; ++c) {
        Chunk after = nxt;
        if (c + 1 < c1) {
            stage_chunk<COND>(sbuf[(c + 1 - c0) & 1], scond[COND ? ((c + 1 - c0) & 1) : 0], nxt, segs, conds,
                              zc_re, zc_im, HAS_IMAG ? 1 : 0);
            if (c + 2 < c1) after = chunks[c + 2];
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const StageBuf &b = sbuf[(c - c0) & 1];
        const uint32_t lo = max(cur.seg_begin, s0), hi = min(cur.seg_begin + cur.seg_count, s1);
        uint32_t s = lo;
        while (s < hi) {
            // batch [s, e): as many segments as fit MB_RUNS class runs
            uint32_t e = s, total = 0;
            while (e < hi) {
                const uint32_t r = b.segs[e - cur.seg_begin].kind_runs & 0xfu;
                if (e > s && total + r > (uint32_t)MB_RUNS) break;
                total += r;
                ++e;
            }
            // phase 1: support test and sign sums of this thread's rows, once for all columns
            uint32_t rb = 0;
            for (uint32_t q = s; q < e; ++q) {
                const Segment &sd = b.segs[q - cur.seg_begin];
                const uint32_t runs = sd.kind_runs & 0xfu, kind = sd.kind_runs >> 30;
                uint32_t need = ALL;
                if (COND) {
                    const SegCond cd = scond[(c - c0) & 1][q - cur.seg_begin];
                    const uint32_t q1 = ((uint32_t)__popc(j_lo & cd.m1) + cd.v) & 1u;
                    const uint32_t q2 = ((uint32_t)__popc(j_lo & cd.m2) + (cd.v >> 1)) & 1u;
                    need = (cd.pat ^ (q1 - 1u)) & ((cd.pat >> 16) ^ (q2 - 1u)) & ALL;
                    s_need[q - cur.seg_begin][tid] = (uint16_t)need;
                }
                if (need) {
                    if (!HAS_IMAG || kind != KIND_IMAG)
                        store_run_sums<Z32>(b.re + sd.rel_begin, runs, sd.cnts, j_lo, base_hi, &a_re[rb][tid]);
                    if (HAS_IMAG && kind != KIND_REAL)
                        store_run_sums<Z32>(b.im + sd.rel_begin, runs, sd.cnts, j_lo, base_hi, &a_im[rb][tid]);
                }
                rb += runs;
            }
            // phase 2: every column gathers its amplitudes and takes the sums from shared memory
            for (int col = 0; col < ncols; ++col) {
                const double2 *__restrict__ in = vin + (size_t)col * (size_t)ld;
                double2 *__restrict__ out = vout + (size_t)col * (size_t)ld;
                double ar[RT], ai[RT];
                if (first_batch) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) ar[r] = ai[r] = 0.0;
                } else {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const double2 o = out[(size_t)(j0 + r * BLOCK)];
                        ar[r] = o.x;
                        ai[r] = o.y;
                    }
                }
                rb = 0;
                for (uint32_t q = s; q < e; ++q) {
                    const Segment &sd = b.segs[q - cur.seg_begin];
                    const uint32_t runs = sd.kind_runs & 0xfu, kind = sd.kind_runs >> 30;
                    const uint32_t need = COND ? (uint32_t)s_need[q - cur.seg_begin][tid] : ALL;
                    if (need) {
                        const uint32_t jx = j0 ^ (sd.x_lo & xmask_local);
                        double2 v[RT];
                        if (!COND || need == ALL)
                            load_tile(in, jx, v);
                        else
                            load_tile_needed(in, jx, need, v);
                        uint32_t ids = sd.ids;
#pragma unroll 1
                        for (uint32_t k = 0; k < runs; ++k) {
                            if (!HAS_IMAG || kind != KIND_IMAG)
                                apply_class_dyn<false>(ids & 0xfu, a_re[rb + k][tid], v, ar, ai);
                            if (HAS_IMAG && kind != KIND_REAL)
                                apply_class_dyn<true>(ids & 0xfu, a_im[rb + k][tid], v, ar, ai);
                            ids >>= 4;
                        }
                    }
                    rb += runs;
                }
#pragma unroll
                for (int r = 0; r < RT; ++r) out[(size_t)(j0 + r * BLOCK)] = make_double2(ar[r], ai[r]);
            }
            first_batch = false;
            s = e;
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
        cur = nxt;
        nxt = after;
    }
}

// Small states (fewer than 2048 local amplitudes): one amplitude per thread, plain sums.
template <int MODE>
__global__ void __launch_bounds__(BLOCK)
k_apply1(const double2 *__restrict__ vin, double2 *__restrict__ vout,
         const Segment *__restrict__ segs, const uint32_t *__restrict__ seg_term_begin,
         uint32_t s0, uint32_t s1,
         const TermZC *__restrict__ zc_re, const TermZC *__restrict__ zc_im, uint32_t dim,
         uint64_t index_base, int accumulate, double2 *__restrict__ partials,
         const SegCond *__restrict__ conds, const double2 *__restrict__ vdot) {
    constexpr bool EXPECT = MODE == 1;
    const uint32_t j = blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = j < dim;
    const uint64_t jz = index_base | j;
    double ar = 0.0, ai = 0.0;
    if (!EXPECT && accumulate && valid) {
        double2 o = vout[j];
        ar = o.x;
        ai = o.y;
    }
    for (uint32_t s = s0; s < s1; ++s) {
        const Segment sd = segs[s];
        const uint32_t kind = sd.kind_runs >> 30;
        const uint32_t tb = seg_term_begin[s], te = tb + ((sd.kind_runs >> 8) & 0xffffu);
        if (conds) {  // support hints: bit 0 of the tile kernel's need mask (this amplitude is "r = 0")
            const SegCond cd = conds[s];
            const uint32_t q1 = ((uint32_t)__popc((uint32_t)jz & cd.m1) + cd.v) & 1u;
            const uint32_t q2 = ((uint32_t)__popc((uint32_t)jz & cd.m2) + (cd.v >> 1)) & 1u;
            if (((cd.pat ^ (q1 - 1u)) & ((cd.pat >> 16) ^ (q2 - 1u)) & 1u) == 0u) continue;
        }
        const double2 v = valid ? vin[j ^ ((uint32_t)sd.x & (dim - 1))] : make_double2(0.0, 0.0);
        double sr = 0.0, si = 0.0;
        for (uint32_t t = tb; t < te; ++t) {
            const TermZC a = zc_re[t];
            const uint32_t par = (uint32_t)__popcll(jz & a.z);
            if (kind != KIND_IMAG) sr += signed_coeff(a.c, par);
            if (kind != KIND_REAL) si += signed_coeff(zc_im[t].c, par);
        }
        ar += sr * v.x - si * v.y;
        ai += sr * v.y + si * v.x;
    }
    if (MODE != 1 && valid) vout[j] = make_double2(ar, ai);
    if (MODE != 0) {
        double er = 0.0, ei = 0.0;
        if (valid) {
            double2 w = MODE == 1 ? vin[j] : vdot[j];
            er = w.x * ar + w.y * ai;
            ei = w.x * ai - w.y * ar;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            er += __shfl_xor_sync(0xffffffffu, er, off);
            ei += __shfl_xor_sync(0xffffffffu, ei, off);
        }
        __shared__ double sm[2 * (BLOCK / 32)];
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            sm[2 * warp] = er;
            sm[2 * warp + 1] = ei;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tr = 0.0, ti = 0.0;
            for (int w = 0; w < BLOCK / 32; ++w) {
                tr += sm[2 * w];
                ti += sm[2 * w + 1];
            }
            partials[blockIdx.x] = make_double2(tr, ti);
        }
    }
}

// K3: diag[j] = sum over the x == 0 group of Re(c_t) (-1)^popc(j & z_t)
template <bool Z32>
__global__ void __launch_bounds__(BLOCK)
k_diagonal(double *__restrict__ out, const TermZC *__restrict__ zc, uint32_t tb, uint32_t te,
           uint64_t dim) {
    const uint64_t j = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (j >= dim) return;
    double s = 0.0;
    for (uint32_t t = tb; t < te; ++t) {
        const TermZC e = zc[t];
        uint32_t par = Z32 ? (uint32_t)__popc((uint32_t)j & (uint32_t)e.z)
                           : (uint32_t)__popcll(j & e.z);
        s += signed_coeff(e.c, par);
    }
    out[j] = s;
}

__global__ void k_zero_c128(double2 *out, uint64_t dim) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < dim) out[j] = make_double2(0.0, 0.0);
}

static inline uint64_t apply_blocks(int local_bits) {
    const uint64_t dim = 1ull << local_bits;
    return local_bits >= TILE_BITS ? dim / TILE : (dim + BLOCK - 1) / BLOCK;
}

template <int MODE>
static int launch_apply_impl(ofb_plan *p, const void *d_in, void *d_out, int local_bits,
                             uint64_t index_base, int64_t g0, int64_t g1, int accumulate,
                             double2 *partials, cudaStream_t s, uint64_t tile0 = 0,
                             uint64_t n_tiles = 0, const void *d_dot = nullptr) {
    if (local_bits > 32) return fail(OFB_ERR_INVALID, "more than 2^32 amplitudes per device");
    const uint64_t dim = 1ull << local_bits;
    uint64_t blocks = apply_blocks(local_bits);
    if (n_tiles) {  // a slab of the tiles (tile kernel only)
        if (local_bits < TILE_BITS || tile0 + n_tiles > blocks) return fail(OFB_ERR_INVALID, "bad tile slab");
        blocks = n_tiles;
    }
    const uint32_t s0 = p->h_seg_begin[g0], s1 = p->h_seg_begin[g1];
    if (s1 <= s0) return fail(OFB_ERR_STATE, "empty segment range");
    const uint32_t c0 = p->h_chunk_of_seg[s0], c1 = p->h_chunk_of_seg[s1 - 1] + 1;
    const double2 *in = (const double2 *)d_in;
    double2 *out = (double2 *)d_out;
    if (local_bits >= TILE_BITS) {
        const uint32_t xmask = (uint32_t)(dim - 1);
#define OFB_GO(ZZ, II, CC)                                                                     \
    k_apply_tile<ZZ, MODE, II, CC><<<(unsigned)blocks, BLOCK, 0, s>>>(                           \
        in, out, p->d_chunks, c0, c1, p->d_segs, s0, s1, p->d_zc, p->d_zc_im, 0, xmask,        \
        index_base, accumulate, partials, p->d_conds, (uint32_t)tile0, (const double2 *)d_dot)
        if (p->z_fits32 && p->has_cond) {  // support hints (ofb_plan_set_support), n <= 32 only
            if (p->has_imag) OFB_GO(true, true, true); else OFB_GO(true, false, true);
        } else if (p->z_fits32) {
            if (p->has_imag) OFB_GO(true, true, false); else OFB_GO(true, false, false);
        } else {
            if (p->has_imag) OFB_GO(false, true, false); else OFB_GO(false, false, false);
        }
#undef OFB_GO
    } else {
        k_apply1<MODE><<<(unsigned)blocks, BLOCK, 0, s>>>(in, out, p->d_segs, p->d_seg_term_begin,
                                                            s0, s1, p->d_zc, p->d_zc_im,
                                                            (uint32_t)dim, index_base, accumulate,
                                                            partials,
                                                            (p->has_cond && p->z_fits32) ? p->d_conds : nullptr,
                                                            (const double2 *)d_dot);
    }
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int launch_apply_slab(ofb_plan *p, const void *d_in, void *d_out, int local_bits, uint64_t tile0,
                      uint64_t n_tiles, cudaStream_t s) {
    return launch_apply_impl<0>(p, d_in, d_out, local_bits, 0, 0, p->n_groups, 0, nullptr, s,
                                tile0, n_tiles);
}

int launch_apply(ofb_plan *p, const void *d_in, void *d_out, int local_bits, uint64_t index_base,
                 int64_t g0, int64_t g1, int accumulate, cudaStream_t s) {
    if (g1 <= g0) {
        if (!accumulate) {
            uint64_t dim = 1ull << local_bits;
            k_zero_c128<<<(unsigned)((dim + 255) / 256), 256, 0, s>>>((double2 *)d_out, dim);
            OFB_LAUNCH_CHECK();
        }
        return OFB_OK;
    }
    return launch_apply_impl<0>(p, d_in, d_out, local_bits, index_base, g0, g1, accumulate,
                                nullptr, s);
}

// H in on a group range plus per-block partial sums of conj(d_dot) . (H in) (complete only when
// this launch finishes the output, i.e. the last launch of an accumulate sequence).
int launch_apply_dot(ofb_plan *p, const void *d_in, void *d_out, int local_bits, uint64_t index_base,
                     int64_t g0, int64_t g1, int accumulate, const void *d_dot, double *d_partials,
                     int64_t *n_partials, cudaStream_t s) {
    if (g1 <= g0) return fail(OFB_ERR_STATE, "empty group range");
    *n_partials = (int64_t)apply_blocks(local_bits);
    return launch_apply_impl<2>(p, d_in, d_out, local_bits, index_base, g0, g1, accumulate,
                                (double2 *)d_partials, s, 0, 0, d_dot);
}
int64_t apply_partial_count(int local_bits) { return (int64_t)apply_blocks(local_bits); }

// K1b: all groups applied to ncols columns `ld` elements apart in one launch (full state only)
int launch_apply_multi(ofb_plan *p, const void *d_in, void *d_out, int ncols, int64_t ld, cudaStream_t s) {
    const int n = p->n_qubits;
    if (n < TILE_BITS || n > 32 || p->n_groups == 0 || ncols < 2)
        return fail(OFB_ERR_STATE, "launch_apply_multi: unsupported configuration");
    const uint64_t dim = 1ull << n;
    const unsigned blocks = (unsigned)(dim / TILE);
    const uint32_t s0 = 0, s1 = p->h_seg_begin[p->n_groups];
    const uint32_t c0 = 0, c1 = p->h_chunk_of_seg[s1 - 1] + 1;
    const uint32_t xmask = (uint32_t)(dim - 1);
    const double2 *in = (const double2 *)d_in;
    double2 *out = (double2 *)d_out;
#define OFB_MULTI(II, CC)                                                                          \
    k_apply_multi<true, II, CC><<<blocks, BLOCK, 0, s>>>(in, out, ncols, ld, p->d_chunks, c0, c1,    \
                                                         p->d_segs, s0, s1, p->d_zc, p->d_zc_im,  \
                                                         xmask, 0ull, p->d_conds)
    if (p->has_cond) {
        if (p->has_imag) OFB_MULTI(true, true); else OFB_MULTI(false, true);
    } else {
        if (p->has_imag) OFB_MULTI(true, false); else OFB_MULTI(false, false);
    }
#undef OFB_MULTI
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
// Opt-in (OFB_MULTI_APPLY=1): measured on B200 the batched kernel is SLOWER per column than the
// single-column one (28-qubit molecular: 1.45x at k = 2, 1.61x at k = 4; profiles/r02_multi_bench.jsonl).
// With the orbit-reduced tables the shared per-term work is under a third of the instructions,
// and in the single-column kernel it doubles as latency hiding for the 16 gathers it follows; the
// batched kernel's second phase has nothing to overlap its gathers with at 8 warps per SM.
// Host -> device copy of a caller's (pageable) vector: above 64 MB through the library's own
// pinned staging and worker threads (hostcopy.cu; a plain pageable cudaMemcpy runs at 11 GB/s and
// the driver serialises several of them), OFB_STAGED_COPY=0 keeps the plain copy.
static inline bool staged_ok(size_t bytes) {
    const char *env = getenv("OFB_STAGED_COPY");
    return bytes >= ((size_t)64 << 20) && !(env && env[0] == '0');
}
static int host_to_device(void *d_dst, const void *h_src, size_t bytes, int device, cudaStream_t s) {
    if (!staged_ok(bytes)) {
        OFB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
        return OFB_OK;
    }
    OFB_CUDA(cudaStreamSynchronize(s));  // the pieces go out on the workers' streams
    return staged_copy(d_dst, const_cast<void *>(h_src), bytes, true, device, 0);
}

static inline bool multi_ok(const ofb_plan *p, int ncols) {
    const char *on = getenv("OFB_MULTI_APPLY");
    return on && on[0] == '1' && ncols >= 2 && p->n_qubits >= TILE_BITS && p->n_qubits <= 32 &&
           p->n_groups > 0;
}

int launch_expectation(ofb_plan *p, const void *d_state, double *h_out, cudaStream_t s) {
    h_out[0] = h_out[1] = 0.0;
    if (p->n_groups == 0) return OFB_OK;
    const int n = p->n_qubits;
    const uint64_t blocks = apply_blocks(n);
    int rc = ensure_partials(p, blocks * sizeof(double2) + 16);
    if (rc) return rc;
    rc = launch_apply_impl<1>(p, d_state, nullptr, n, 0, 0, p->n_groups, 0,
                                 (double2 *)p->d_partials, s);
    if (rc) return rc;
    double *d_res = p->d_partials + 2 * blocks;
    rc = reduce_partials(p->d_partials, (int64_t)blocks, 2, d_res, s);
    if (rc) return rc;
    OFB_CUDA(cudaMemcpyAsync(p->h_pinned_scalars, d_res, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    h_out[0] = p->h_pinned_scalars[0];
    h_out[1] = p->h_pinned_scalars[1];
    return OFB_OK;
}

int launch_diagonal(ofb_plan *p, double *d_out, cudaStream_t s) {
    const uint64_t dim = 1ull << p->n_qubits;
    const uint64_t blocks = (dim + BLOCK - 1) / BLOCK;
    if (blocks > 0x7fffffffull) return fail(OFB_ERR_INVALID, "state too large for one launch");
    uint32_t te = (uint32_t)p->n_diag_terms;
    if (p->z_fits32)
        k_diagonal<true><<<(unsigned)blocks, BLOCK, 0, s>>>(d_out, p->d_zc, 0, te, dim);
    else
        k_diagonal<false><<<(unsigned)blocks, BLOCK, 0, s>>>(d_out, p->d_zc, 0, te, dim);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_apply_groups(ofb_plan *p, const void *d_in, void *d_out, int local_bits,
                     uint64_t index_base, int64_t g_begin, int64_t g_end, int accumulate,
                     void *stream) {
    if (!p || !d_in || !d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    if (local_bits < 0 || local_bits > p->n_qubits) return fail(OFB_ERR_INVALID, "bad local_bits");
    if (g_begin < 0 || g_end > p->n_groups || g_begin > g_end)
        return fail(OFB_ERR_INVALID, "bad group range");
    if (d_in == d_out) return fail(OFB_ERR_INVALID, "d_in and d_out must not alias");
    if (local_bits < 64 && (index_base & ((1ull << local_bits) - 1)))
        return fail(OFB_ERR_INVALID, "index_base overlaps the local index bits");
    OFB_CUDA(cudaSetDevice(p->device));
    return launch_apply(p, d_in, d_out, local_bits, index_base, g_begin, g_end, accumulate,
                        (cudaStream_t)stream);
}

int ofb_apply(ofb_plan *p, const void *d_in, void *d_out, int ncols, int64_t ld, void *stream) {
    if (!p || !d_in || !d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    const int64_t dim = 1ll << p->n_qubits;
    if (ncols < 0 || (ncols > 1 && ld < dim)) return fail(OFB_ERR_INVALID, "bad ncols/ld");
    if (d_in == d_out) return fail(OFB_ERR_INVALID, "d_in and d_out must not alias");
    OFB_CUDA(cudaSetDevice(p->device));
    if (multi_ok(p, ncols))  // K1b: the per-term work is shared by the columns
        return launch_apply_multi(p, d_in, d_out, ncols, ld, (cudaStream_t)stream);
    for (int c = 0; c < ncols; ++c) {
        const char *in = (const char *)d_in + (size_t)c * ld * 16;
        char *out = (char *)d_out + (size_t)c * ld * 16;
        int rc = launch_apply(p, in, out, p->n_qubits, 0, 0, p->n_groups, 0, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return OFB_OK;
}

int ofb_apply_host(ofb_plan *p, const void *h_in, void *h_out, int ncols) {
    if (!p || !h_in || !h_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    if (ncols < 0) return fail(OFB_ERR_INVALID, "bad ncols");
    OFB_CUDA(cudaSetDevice(p->device));
    const size_t dim = (size_t)1 << p->n_qubits;
    const size_t col_bytes = dim * 16;
    cudaStream_t s = 0;
    int rc;
    if (multi_ok(p, ncols)) {
        // matmat: batches of up to 8 columns through K1b (as many as fit a quarter of the free
        // device memory, in and out staged together)
        size_t free_b = 0, total_b = 0;
        OFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += p->stage_in_bytes + p->stage_out_bytes;
        int batch = (int)std::min<size_t>(8, std::max<size_t>(1, free_b / 4 / (2 * col_bytes)));
        batch = std::min(batch, ncols);
        if ((rc = ensure_stage(p, (size_t)batch * col_bytes, (size_t)batch * col_bytes))) return rc;
        for (int c = 0; c < ncols; c += batch) {
            const int nb = std::min(batch, ncols - c);
            OFB_CUDA(cudaMemcpyAsync(p->d_stage_in, (const char *)h_in + (size_t)c * col_bytes,
                                     (size_t)nb * col_bytes, cudaMemcpyHostToDevice, s));
            if (nb >= 2)
                rc = launch_apply_multi(p, p->d_stage_in, p->d_stage_out, nb, (int64_t)dim, s);
            else
                rc = launch_apply(p, p->d_stage_in, p->d_stage_out, p->n_qubits, 0, 0, p->n_groups, 0, s);
            if (rc) return rc;
            OFB_CUDA(cudaMemcpyAsync((char *)h_out + (size_t)c * col_bytes, p->d_stage_out,
                                     (size_t)nb * col_bytes, cudaMemcpyDeviceToHost, s));
        }
        OFB_CUDA(cudaStreamSynchronize(s));
        return OFB_OK;
    }
    rc = ensure_stage(p, col_bytes, col_bytes);
    if (rc) return rc;
    // Large states: the output goes back slab by slab on a second stream while the remaining
    // tiles are still being computed (the input must be complete before the first tile starts:
    // a gather reaches anywhere).  Host time of the call = H2D + kernel + the last slab's D2H.
    const uint64_t tiles = p->n_qubits >= TILE_BITS ? dim / TILE : 0;
    const int slabs = (p->n_groups > 0 && tiles >= 8 * 1024) ? HOST_SLABS : 1;
    if (slabs > 1 && !p->copy_stream) {
        OFB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < HOST_SLABS; ++k)
            OFB_CUDA(cudaEventCreateWithFlags(&p->slab_done[k], cudaEventDisableTiming));
    }
    // Groups whose x mask has no bit among the top log2(slabs) index bits gather inside their own
    // slab: their share of the work (62 % of the 28-qubit molecular table) starts as soon as that
    // slab of the INPUT has arrived, while the worker threads copy the next slab; the remaining
    // groups follow as accumulate launches once the whole input is on the device.
    int64_t g_local = 0;
    if (slabs > 1 && staged_ok(col_bytes / slabs) && !(getenv("OFB_H2D_OVERLAP") && getenv("OFB_H2D_OVERLAP")[0] == '0')) {
        const uint64_t limit = (uint64_t)dim / (uint64_t)slabs;
        g_local = std::lower_bound(p->h_group_x.begin(), p->h_group_x.end(), limit) - p->h_group_x.begin();
    }
    for (int c = 0; c < ncols; ++c) {
        const char *src_col = (const char *)h_in + c * col_bytes;
        if (g_local == 0) {
            rc = host_to_device(p->d_stage_in, src_col, col_bytes, p->device, s);
            if (rc) return rc;
        }
        if (slabs == 1) {
            rc = launch_apply(p, p->d_stage_in, p->d_stage_out, p->n_qubits, 0, 0, p->n_groups, 0, s);
            if (rc) return rc;
            if (staged_ok(col_bytes)) {
                OFB_CUDA(cudaStreamSynchronize(s));
                if ((rc = staged_copy(p->d_stage_out, (char *)h_out + c * col_bytes, col_bytes, false, p->device, 0)))
                    return rc;
            } else {
                OFB_CUDA(cudaMemcpyAsync((char *)h_out + c * col_bytes, p->d_stage_out, col_bytes,
                                         cudaMemcpyDeviceToHost, s));
            }
            continue;
        }
        const uint64_t per = tiles / slabs;
        if (g_local > 0) {
            OFB_CUDA(cudaStreamSynchronize(s));  // the previous column's kernels are done with stage_in
            const size_t in_slab = col_bytes / slabs;
            for (int k = 0; k < slabs; ++k) {
                rc = staged_copy((char *)p->d_stage_in + k * in_slab, const_cast<char *>(src_col) + k * in_slab, in_slab,
                                 true, p->device, 0);
                if (!rc)
                    rc = launch_apply_impl<0>(p, p->d_stage_in, p->d_stage_out, p->n_qubits, 0, 0, g_local, 0, nullptr,
                                              s, k * per, per);
                if (rc) return rc;
                if (g_local == p->n_groups) OFB_CUDA(cudaEventRecord(p->slab_done[k], s));  // slab complete
            }
        }
        for (int k = 0; k < slabs && !(g_local > 0 && g_local == p->n_groups); ++k) {
            // all launches are queued before the first output copy is issued
            if (g_local > 0)
                rc = launch_apply_impl<0>(p, p->d_stage_in, p->d_stage_out, p->n_qubits, 0, g_local, p->n_groups, 1,
                                          nullptr, s, k * per, per);
            else
                rc = launch_apply_slab(p, p->d_stage_in, p->d_stage_out, p->n_qubits, k * per, per, s);
            if (rc) return rc;
            OFB_CUDA(cudaEventRecord(p->slab_done[k], s));
        }
        const size_t slab_bytes = col_bytes / slabs;
        for (int k = 0; k < slabs; ++k) {
            char *dst = (char *)h_out + c * col_bytes + k * slab_bytes;
            const char *src = (const char *)p->d_stage_out + k * slab_bytes;
            if (staged_ok(slab_bytes)) {  // worker threads fault the caller's pages in parallel
                OFB_CUDA(cudaEventSynchronize(p->slab_done[k]));
                if ((rc = staged_copy(const_cast<char *>(src), dst, slab_bytes, false, p->device, 0))) return rc;
            } else {
                OFB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->slab_done[k], 0));
                OFB_CUDA(cudaMemcpyAsync(dst, src, slab_bytes, cudaMemcpyDeviceToHost, p->copy_stream));
            }
        }
        OFB_CUDA(cudaStreamSynchronize(p->copy_stream));  // stage_out is free for the next column
    }
    OFB_CUDA(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_expectation(ofb_plan *p, const void *d_state, double *h_out_ri, void *stream) {
    if (!p || !d_state || !h_out_ri) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    return launch_expectation(p, d_state, h_out_ri, (cudaStream_t)stream);
}

int ofb_expectation_host(ofb_plan *p, const void *h_state, double *h_out_ri) {
    if (!p || !h_state || !h_out_ri) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    const size_t bytes = ((size_t)1 << p->n_qubits) * 16;
    int rc = ensure_stage(p, bytes, 0);
    if (rc) return rc;
    rc = host_to_device(p->d_stage_in, h_state, bytes, p->device, 0);
    if (rc) return rc;
    return launch_expectation(p, p->d_stage_in, h_out_ri, 0);
}

int ofb_diagonal(ofb_plan *p, double *d_out, void *stream) {
    if (!p || !d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    return launch_diagonal(p, d_out, (cudaStream_t)stream);
}

int ofb_diagonal_host(ofb_plan *p, double *h_out) {
    if (!p || !h_out) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    OFB_CUDA(cudaSetDevice(p->device));
    const size_t bytes = ((size_t)1 << p->n_qubits) * 8;
    int rc = ensure_stage(p, 0, bytes);
    if (rc) return rc;
    rc = launch_diagonal(p, (double *)p->d_stage_out, 0);
    if (rc) return rc;
    if (staged_ok(bytes)) {
        OFB_CUDA(cudaStreamSynchronize(0));
        return staged_copy(p->d_stage_out, h_out, bytes, false, p->device, 0);
    }
    OFB_CUDA(cudaMemcpyAsync(h_out, p->d_stage_out, bytes, cudaMemcpyDeviceToHost, 0));
    OFB_CUDA(cudaStreamSynchronize(0));
    return OFB_OK;
}

}  // extern "C"
