// K1 matrix-free H|psi>, K2 fused expectation, K3 diagonal.
//
// Gather form:
//     out[j] = sum_g S_g(j) * in[j ^ x_g],   S_g(j) = sum_{t in g} c_t (-1)^popc(j & z_t)
// Each thread owns R amplitudes j_r = tile + r*BLOCK + tid (stride BLOCK so that every
// 128-bit load/store instruction of a warp covers 512 contiguous bytes; XOR with x only
// permutes lanes inside aligned blocks, so loads stay sector-coalesced for any mask).
// The parity popc(j & z) is evaluated once per (thread, term) for r = 0; the other r differ
// only in index bits log2(BLOCK).. and flip the sign by z's bits there, which is folded
// into the high word of the coefficient with one LOP3 per (term, amplitude).  The running
// S_g is a real scalar for groups whose phased coefficients are all real (or all
// imaginary), so the common Hermitian/real case costs one DADD per term-amplitude and
// 2 DFMA per (group, amplitude).  Term tables are read with warp-uniform 128-bit loads.
#include "ofb_internal.cuh"

namespace ofb {

constexpr int BLOCK = 256;
constexpr int BLOCK_BITS = 8;

__device__ __forceinline__ double flip_sign_hi(uint32_t hi, uint32_t lo) {
    return __hiloint2double((int)hi, (int)lo);
}

template <int R, bool Z32>
struct TileIdx {
    uint64_t j0;  // local index of amplitude r = 0
    uint64_t jz;  // global index (index_base | j0) for the sign parity
};

// sign bit (at bit 31) of amplitude r relative to amplitude 0: parity(r & (z >> BLOCK_BITS))
template <int R>
__device__ __forceinline__ void r_flips(uint64_t z, uint32_t (&flip)[R]) {
    uint32_t zb = (uint32_t)(z >> BLOCK_BITS);
    flip[0] = 0;
#pragma unroll
    for (int r = 1; r < R; ++r) {
        // highest set bit k of r: flip[r] = flip[r - 2^k] ^ bit k of zb
        int k = 31 - __builtin_clz((unsigned)r);
        flip[r] = flip[r - (1 << k)] ^ ((zb << (31 - k)) & 0x80000000u);
    }
}

template <int R, bool Z32, bool EXPECT>
__global__ void __launch_bounds__(BLOCK)
k_apply(const double2 *__restrict__ vin, double2 *__restrict__ vout,
        const GroupDesc *__restrict__ groups, uint32_t g0, uint32_t g1,
        const TermZC *__restrict__ zc, const double *__restrict__ ci, uint64_t dim,
        uint64_t xmask_local, uint64_t index_base, int accumulate, double2 *__restrict__ partials) {
    const uint64_t tile = (uint64_t)blockIdx.x * (uint64_t)(BLOCK * R);
    const uint64_t j0 = tile + threadIdx.x;
    const uint64_t jz = index_base | j0;
    const uint32_t jz32 = (uint32_t)jz;

    double ar[R], ai[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint64_t j = j0 + (uint64_t)r * BLOCK;
        valid[r] = j < dim;
        ar[r] = 0.0;
        ai[r] = 0.0;
        if (!EXPECT && accumulate && valid[r]) {
            double2 o = vout[j];
            ar[r] = o.x;
            ai[r] = o.y;
        }
    }

    for (uint32_t g = g0; g < g1; ++g) {
        const GroupDesc gd = groups[g];
        const uint64_t xl = gd.x & xmask_local;
        const uint32_t kind = gd.count >> 30;
        const uint32_t tb = gd.begin, te = gd.begin + (gd.count & 0x3fffffffu);
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint64_t j = (j0 + (uint64_t)r * BLOCK) ^ xl;
            v[r] = valid[r] ? vin[j] : make_double2(0.0, 0.0);
        }
        if (kind != KIND_COMPLEX) {
            double s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) s[r] = 0.0;
#pragma unroll 2
            for (uint32_t t = tb; t < te; ++t) {
                const TermZC e = zc[t];
                uint32_t par = Z32 ? (uint32_t)__popc(jz32 & (uint32_t)e.z)
                                   : (uint32_t)__popcll(jz & e.z);
                uint32_t hi = (uint32_t)__double2hiint(e.c) ^ (par << 31);
                uint32_t lo = (uint32_t)__double2loint(e.c);
                uint32_t flip[R];
                r_flips<R>(e.z, flip);
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] += flip_sign_hi(hi ^ flip[r], lo);
            }
            if (kind == KIND_REAL) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ar[r] = fma(s[r], v[r].x, ar[r]);
                    ai[r] = fma(s[r], v[r].y, ai[r]);
                }
            } else {  // (i s) * (vx + i vy) = -s vy + i s vx
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ar[r] = fma(-s[r], v[r].y, ar[r]);
                    ai[r] = fma(s[r], v[r].x, ai[r]);
                }
            }
        } else {
            double sr[R], si[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                sr[r] = 0.0;
                si[r] = 0.0;
            }
            for (uint32_t t = tb; t < te; ++t) {
                const TermZC e = zc[t];
                const double cim = ci[t];
                uint32_t par = Z32 ? (uint32_t)__popc(jz32 & (uint32_t)e.z)
                                   : (uint32_t)__popcll(jz & e.z);
                uint32_t sgn = par << 31;
                uint32_t hr = (uint32_t)__double2hiint(e.c) ^ sgn, lr = (uint32_t)__double2loint(e.c);
                uint32_t hi = (uint32_t)__double2hiint(cim) ^ sgn, li = (uint32_t)__double2loint(cim);
                uint32_t flip[R];
                r_flips<R>(e.z, flip);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    sr[r] += flip_sign_hi(hr ^ flip[r], lr);
                    si[r] += flip_sign_hi(hi ^ flip[r], li);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ar[r] = fma(sr[r], v[r].x, ar[r]);
                ar[r] = fma(-si[r], v[r].y, ar[r]);
                ai[r] = fma(sr[r], v[r].y, ai[r]);
                ai[r] = fma(si[r], v[r].x, ai[r]);
            }
        }
    }

    if (!EXPECT) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (valid[r]) vout[j0 + (uint64_t)r * BLOCK] = make_double2(ar[r], ai[r]);
    } else {
        // conj(in[j]) * (H in)[j], block-reduced in a fixed order (deterministic)
        double er = 0.0, ei = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (valid[r]) {
                double2 w = vin[j0 + (uint64_t)r * BLOCK];
                er += w.x * ar[r] + w.y * ai[r];
                ei += w.x * ai[r] - w.y * ar[r];
            }
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
        s += flip_sign_hi((uint32_t)__double2hiint(e.c) ^ (par << 31), (uint32_t)__double2loint(e.c));
    }
    out[j] = s;
}

__global__ void k_zero_c128(double2 *out, uint64_t dim) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < dim) out[j] = make_double2(0.0, 0.0);
}

static int pick_r(int local_bits) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("OFB_APPLY_R");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) {
        return (local_bits >= BLOCK_BITS + 3) ? forced : 1;
    }
    return local_bits >= 14 ? 4 : 1;
}

template <bool EXPECT>
static int launch_apply_impl(ofb_plan *p, const void *d_in, void *d_out, int local_bits,
                             uint64_t index_base, int64_t g0, int64_t g1, int accumulate,
                             double2 *partials, int64_t *n_blocks_out, cudaStream_t s) {
    const uint64_t dim = 1ull << local_bits;
    const uint64_t xmask_local = dim - 1;
    const int R = pick_r(local_bits);
    const uint64_t per_block = (uint64_t)BLOCK * R;
    const uint64_t blocks = (dim + per_block - 1) / per_block;
    if (blocks > 0x7fffffffull) return fail(OFB_ERR_INVALID, "state too large for one launch");
    if (n_blocks_out) *n_blocks_out = (int64_t)blocks;
    const double2 *in = (const double2 *)d_in;
    double2 *out = (double2 *)d_out;
    const bool z32 = p->z_fits32;
#define OFB_GO(RR, ZZ)                                                                      \
    k_apply<RR, ZZ, EXPECT><<<(unsigned)blocks, BLOCK, 0, s>>>(                             \
        in, out, p->d_groups, (uint32_t)g0, (uint32_t)g1, p->d_zc, p->d_ci, dim, xmask_local, \
        index_base, accumulate, partials)
    switch (R) {
        case 1: if (z32) OFB_GO(1, true); else OFB_GO(1, false); break;
        case 2: if (z32) OFB_GO(2, true); else OFB_GO(2, false); break;
        case 4: if (z32) OFB_GO(4, true); else OFB_GO(4, false); break;
        default: if (z32) OFB_GO(8, true); else OFB_GO(8, false); break;
    }
#undef OFB_GO
    OFB_LAUNCH_CHECK();
    return OFB_OK;
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
    return launch_apply_impl<false>(p, d_in, d_out, local_bits, index_base, g0, g1, accumulate,
                                    nullptr, nullptr, s);
}

int launch_expectation(ofb_plan *p, const void *d_state, double *h_out, cudaStream_t s) {
    h_out[0] = h_out[1] = 0.0;
    if (p->n_groups == 0) return OFB_OK;
    const int n = p->n_qubits;
    const uint64_t dim = 1ull << n;
    const int R = pick_r(n);
    const uint64_t blocks = (dim + (uint64_t)BLOCK * R - 1) / ((uint64_t)BLOCK * R);
    int rc = ensure_partials(p, blocks * sizeof(double2) + 16);
    if (rc) return rc;
    int64_t nb = 0;
    rc = launch_apply_impl<true>(p, d_state, nullptr, n, 0, 0, p->n_groups, 0,
                                 (double2 *)p->d_partials, &nb, s);
    if (rc) return rc;
    double *d_res = p->d_partials + 2 * blocks;
    rc = reduce_partials(p->d_partials, nb, 2, d_res, s);
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
    int rc = ensure_stage(p, col_bytes, col_bytes);
    if (rc) return rc;
    cudaStream_t s = 0;
    for (int c = 0; c < ncols; ++c) {
        OFB_CUDA(cudaMemcpyAsync(p->d_stage_in, (const char *)h_in + c * col_bytes, col_bytes,
                                 cudaMemcpyHostToDevice, s));
        rc = launch_apply(p, p->d_stage_in, p->d_stage_out, p->n_qubits, 0, 0, p->n_groups, 0, s);
        if (rc) return rc;
        OFB_CUDA(cudaMemcpyAsync((char *)h_out + c * col_bytes, p->d_stage_out, col_bytes,
                                 cudaMemcpyDeviceToHost, s));
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
    OFB_CUDA(cudaMemcpyAsync(p->d_stage_in, h_state, bytes, cudaMemcpyHostToDevice, 0));
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
    OFB_CUDA(cudaMemcpyAsync(h_out, p->d_stage_out, bytes, cudaMemcpyDeviceToHost, 0));
    OFB_CUDA(cudaStreamSynchronize(0));
    return OFB_OK;
}

}  // extern "C"
