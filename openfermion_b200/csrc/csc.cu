// K4: two-pass CSC materialisation (count -> scan -> fill), replacing the reference's
// per-term kron / sparse-product chain and scipy's coo->csc + sum_duplicates +
// eliminate_zeros (linalg/sparse_tools.py:76-194).
//
// One thread owns one column c.  Scatter rows are grouped by x-mask in ascending x; the
// entry of group g in column c sits at row c ^ x_g with value
//     V_g(c) = sum over rows t of g with (c & occ_mask_t) == occ_val_t of c_t (-1)^popc(c & z_t)
// summed in term order (scipy sums duplicates left to right after a per-column sort that
// is stable for <= 16 entries).  Exact zeros are dropped (eliminate_zeros).  Rows inside a
// column must ascend: with the sibling-subtree table of the x trie (plan.cu) the rank of
// group g among all groups of column c is  sum_{b in bits(c ^ x_g)} sib[g][b]; the count
// pass records the non-zero groups as a rank-ordered bitmap per column plus per-word
// prefix counts, so the fill pass places every entry directly at its sorted position.
// All loops over groups/rows are warp-uniform (table reads are broadcast loads).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ofb_internal.cuh"

namespace ofb {

constexpr int CB = 256;

template <bool Z32>
__device__ __forceinline__ void group_value(const ScatterRow *__restrict__ rows, uint32_t b,
                                            uint32_t e, uint64_t c, int n, double &sr, double &si) {
    sr = 0.0;
    si = 0.0;
    bool first = true;  // the sum starts from its first element: 0 + (-0) would lose a signed zero
    for (uint32_t k = b; k < e; ++k) {
        const ScatterRow r = rows[k];
        // (a zero coefficient leaves no triplet in the reference: coo_matrix(0) is empty)
        if ((c & r.occ_mask) == r.occ_val && (r.cr != 0.0 || r.ci != 0.0)) {
            const double2 v = scatter_value(r, c, n);
            sr = first ? v.x : sr + v.x;
            si = first ? v.y : si + v.y;
            first = false;
        }
    }
}

__device__ __forceinline__ uint32_t rank_of(const uint32_t *__restrict__ sib, uint32_t g, int n,
                                            uint64_t row_bits) {
    uint32_t rank = 0;
    const uint32_t *s = sib + (size_t)g * n;
    while (row_bits) {
        int b = __ffsll((long long)row_bits) - 1;
        rank += s[b];
        row_bits &= row_bits - 1;
    }
    return rank;
}

template <bool Z32>
__global__ void __launch_bounds__(CB)
k_csc_count(const ScatterRow *__restrict__ rows, const GroupDesc *__restrict__ groups, uint32_t G,
            const uint32_t *__restrict__ sib, int n, uint64_t dim, uint32_t words,
            uint32_t *__restrict__ bitmap, uint32_t *__restrict__ prefix,
            int64_t *__restrict__ colcount) {
    const uint64_t c = (uint64_t)blockIdx.x * CB + threadIdx.x;
    if (c >= dim) return;
    for (uint32_t w = 0; w < words; ++w) bitmap[(size_t)w * dim + c] = 0;
    for (uint32_t g = 0; g < G; ++g) {
        const GroupDesc gd = groups[g];
        double sr, si;
        group_value<Z32>(rows, gd.begin, gd.begin + (gd.count & 0x3fffffffu), c, n, sr, si);
        if (sr != 0.0 || si != 0.0) {
            uint32_t rank = rank_of(sib, g, n, c ^ gd.x);
            bitmap[(size_t)(rank >> 5) * dim + c] |= 1u << (rank & 31);
        }
    }
    uint32_t run = 0;
    for (uint32_t w = 0; w < words; ++w) {
        prefix[(size_t)w * dim + c] = run;
        run += __popc(bitmap[(size_t)w * dim + c]);
    }
    colcount[c] = run;
}

template <bool Z32, typename I>
__global__ void __launch_bounds__(CB)
k_csc_fill(const ScatterRow *__restrict__ rows, const GroupDesc *__restrict__ groups, uint32_t G,
           const uint32_t *__restrict__ sib, int n, uint64_t dim,
           const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ prefix,
           const int64_t *__restrict__ colptr, I *__restrict__ indices, double2 *__restrict__ data) {
    const uint64_t c = (uint64_t)blockIdx.x * CB + threadIdx.x;
    if (c >= dim) return;
    const int64_t base = colptr[c];
    for (uint32_t g = 0; g < G; ++g) {
        const GroupDesc gd = groups[g];
        double sr, si;
        group_value<Z32>(rows, gd.begin, gd.begin + (gd.count & 0x3fffffffu), c, n, sr, si);
        if (sr != 0.0 || si != 0.0) {
            const uint64_t row = c ^ gd.x;
            uint32_t rank = rank_of(sib, g, n, row);
            size_t at = (size_t)(rank >> 5) * dim + c;
            uint32_t below = bitmap[at] & ((1u << (rank & 31)) - 1u);
            int64_t pos = base + prefix[at] + __popc(below);
            indices[pos] = (I)row;
            data[pos] = make_double2(sr, si);
        }
    }
}

// ---- exclusive scan of int64 counts (three small kernels) -------------------
constexpr int SB = 1024;
__global__ void __launch_bounds__(SB) k_scan_local(const int64_t *__restrict__ in, int64_t n,
                                                   int64_t *__restrict__ out,
                                                   int64_t *__restrict__ block_sums) {
    __shared__ int64_t warp_tot[SB / 32];
    int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x;
    int64_t v = i < n ? in[i] : 0;
    int64_t incl = v;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t t = warp_tot[lane];
        int64_t ti = t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int64_t o = __shfl_up_sync(0xffffffffu, ti, off);
            if (lane >= off) ti += o;
        }
        warp_tot[lane] = ti - t;  // exclusive warp offsets
        if (lane == 31) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    if (i < n) out[i] = incl - v + warp_tot[warp];
}
__global__ void k_scan_sums(int64_t *block_sums, int64_t nb, int64_t *total) {
    // single thread: nb <= 2^n / 1024, tiny next to the count pass
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int64_t b = 0; b < nb; ++b) {
            int64_t t = block_sums[b];
            block_sums[b] = run;
            run += t;
        }
        *total = run;
    }
}
__global__ void __launch_bounds__(SB) k_scan_add(int64_t *out, int64_t n,
                                                 const int64_t *__restrict__ block_sums) {
    int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
}

template <typename I>
__global__ void k_convert_ptr(const int64_t *__restrict__ in, int64_t n, I *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (I)in[i];
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_csc_count(ofb_plan *p, int64_t *nnz, int *index_bits, void *stream) {
    if (!p || !nnz) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->n_qubits > 30) return fail(OFB_ERR_INVALID, "CSC materialisation supports n_qubits <= 30");
    OFB_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t dim = 1ull << p->n_qubits;
    const uint32_t G = (uint32_t)p->n_groups;
    const uint32_t words = (G + 31) / 32;
    cudaFree(p->d_colptr);
    cudaFree(p->d_bitmap);
    p->d_colptr = nullptr;
    p->d_bitmap = nullptr;
    p->csc_nnz = -1;
    const int64_t nb = (int64_t)((dim + SB - 1) / SB);
    // colptr[dim + 1] then counts[dim] then block sums[nb]
    OFB_CUDA(cudaMalloc(&p->d_colptr, (2 * dim + 1 + nb + 1) * sizeof(int64_t)));
    OFB_CUDA(cudaMalloc(&p->d_bitmap, p->csc_mode == 1 ? 16 : (size_t)2 * std::max<uint32_t>(words, 1) * dim * sizeof(uint32_t)));
    p->bitmap_words = words;
    int64_t *colptr = p->d_colptr, *counts = p->d_colptr + dim + 1, *bsums = counts + dim;
    uint32_t *bitmap = p->d_bitmap, *prefix = p->d_bitmap + (size_t)words * dim;
    // group descriptors for projected plans are not uploaded by plan_create: build on demand
    if (!p->d_groups) {
        OFB_TRY
        std::vector<GroupDesc> groups(p->n_groups);
        for (int64_t g = 0; g < p->n_groups; ++g)
            groups[g] = GroupDesc{p->h_group_x[g], p->h_group_begin[g],
                                  p->h_group_begin[g + 1] - p->h_group_begin[g]};
        OFB_CUDA(cudaMalloc(&p->d_groups, std::max<size_t>(groups.size(), 1) * sizeof(GroupDesc)));
        if (!groups.empty())
            OFB_CUDA(cudaMemcpy(p->d_groups, groups.data(), groups.size() * sizeof(GroupDesc),
                                cudaMemcpyHostToDevice));
        OFB_CATCH((void)0)
    }
    if (p->csc_mode == 1) {
        int rc = csc_exact_count(p, counts, s);
        if (rc) return rc;
    } else if (csc_tile_ok(p) && !getenv("OFB_CSC_COLUMN_KERNEL")) {
        int rc = csc_tile_count(p, words, bitmap, prefix, counts, s);  // K4 v2 (csc_tile.cu)
        if (rc) return rc;
    } else {
        const unsigned grid = (unsigned)((dim + CB - 1) / CB);
        if (p->z_fits32)
            k_csc_count<true><<<grid, CB, 0, s>>>(p->d_rows, p->d_groups, G, p->d_sib, p->n_qubits,
                                                  dim, words, bitmap, prefix, counts);
        else
            k_csc_count<false><<<grid, CB, 0, s>>>(p->d_rows, p->d_groups, G, p->d_sib, p->n_qubits,
                                                   dim, words, bitmap, prefix, counts);
        OFB_LAUNCH_CHECK();
    }
    k_scan_local<<<(unsigned)nb, SB, 0, s>>>(counts, (int64_t)dim, colptr, bsums);
    OFB_LAUNCH_CHECK();
    k_scan_sums<<<1, 32, 0, s>>>(bsums, nb, colptr + dim);
    OFB_LAUNCH_CHECK();
    k_scan_add<<<(unsigned)nb, SB, 0, s>>>(colptr, (int64_t)dim, bsums);
    OFB_LAUNCH_CHECK();
    OFB_CUDA(cudaMemcpyAsync(p->h_pinned_scalars, colptr + dim, sizeof(int64_t),
                             cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    int64_t total;
    memcpy(&total, p->h_pinned_scalars, sizeof(int64_t));
    p->csc_nnz = total;
    *nnz = total;
    // scipy get_index_dtype: int32 iff max(nnz, dim) < 2^31
    if (index_bits) *index_bits = (std::max<int64_t>(total, (int64_t)dim) < (1ll << 31)) ? 32 : 64;
    return OFB_OK;
}

int ofb_csc_fill(ofb_plan *p, void *d_indptr, void *d_indices, void *d_data, int index_bits,
                 void *stream) {
    if (!p || !d_indptr) return fail(OFB_ERR_INVALID, "NULL argument");
    if (index_bits != 32 && index_bits != 64) return fail(OFB_ERR_INVALID, "index_bits must be 32 or 64");
    if (p->csc_nnz < 0 || !p->d_colptr) return fail(OFB_ERR_STATE, "ofb_csc_count must run first");
    if (p->csc_nnz > 0 && (!d_indices || !d_data)) return fail(OFB_ERR_INVALID, "NULL output");
    OFB_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t dim = 1ull << p->n_qubits;
    const uint32_t G = (uint32_t)p->n_groups;
    const uint32_t words = (uint32_t)p->bitmap_words;
    const uint32_t *bitmap = p->d_bitmap, *prefix = p->d_bitmap + (size_t)words * dim;
    const unsigned grid = (unsigned)((dim + CB - 1) / CB);
    const unsigned pgrid = (unsigned)((dim + 1 + 255) / 256);
    if (p->csc_mode == 0 && p->csc_nnz > 0 && csc_tile_ok(p) && !getenv("OFB_CSC_COLUMN_KERNEL")) {
        if (index_bits == 32)
            k_convert_ptr<int32_t><<<pgrid, 256, 0, s>>>(p->d_colptr, (int64_t)dim + 1, (int32_t *)d_indptr);
        else
            k_convert_ptr<int64_t><<<pgrid, 256, 0, s>>>(p->d_colptr, (int64_t)dim + 1, (int64_t *)d_indptr);
        OFB_LAUNCH_CHECK();
        return csc_tile_fill(p, words, p->d_bitmap, p->d_bitmap + (size_t)words * dim, d_indptr,
                             d_indices, d_data, index_bits, s);
    }
#define OFB_FILL(ZZ, II)                                                                       \
    k_csc_fill<ZZ, II><<<grid, CB, 0, s>>>(p->d_rows, p->d_groups, G, p->d_sib, p->n_qubits, dim, \
                                           bitmap, prefix, p->d_colptr, (II *)d_indices,        \
                                           (double2 *)d_data)
    if (index_bits == 32) {
        k_convert_ptr<int32_t><<<pgrid, 256, 0, s>>>(p->d_colptr, (int64_t)dim + 1, (int32_t *)d_indptr);
        OFB_LAUNCH_CHECK();
        if (p->csc_nnz > 0 && p->csc_mode == 1) {
            int rc = csc_exact_fill(p, p->d_colptr, d_indices, d_data, 32, s);
            if (rc) return rc;
        } else if (p->csc_nnz > 0) {
            if (p->z_fits32) OFB_FILL(true, int32_t); else OFB_FILL(false, int32_t);
            OFB_LAUNCH_CHECK();
        }
    } else {
        k_convert_ptr<int64_t><<<pgrid, 256, 0, s>>>(p->d_colptr, (int64_t)dim + 1, (int64_t *)d_indptr);
        OFB_LAUNCH_CHECK();
        if (p->csc_nnz > 0 && p->csc_mode == 1) {
            int rc = csc_exact_fill(p, p->d_colptr, d_indices, d_data, 64, s);
            if (rc) return rc;
        } else if (p->csc_nnz > 0) {
            if (p->z_fits32) OFB_FILL(true, int64_t); else OFB_FILL(false, int64_t);
            OFB_LAUNCH_CHECK();
        }
    }
#undef OFB_FILL
    return OFB_OK;
}

int ofb_csc_fill_host(ofb_plan *p, void *h_indptr, void *h_indices, void *h_data, int index_bits) {
    if (!p || !h_indptr) return fail(OFB_ERR_INVALID, "NULL argument");
    if (index_bits != 32 && index_bits != 64) return fail(OFB_ERR_INVALID, "index_bits must be 32 or 64");
    if (p->csc_nnz < 0) return fail(OFB_ERR_STATE, "ofb_csc_count must run first");
    OFB_CUDA(cudaSetDevice(p->device));
    const size_t dim = (size_t)1 << p->n_qubits;
    const size_t ib = index_bits / 8;
    const size_t nnz = (size_t)p->csc_nnz;
    void *d_indptr = nullptr, *d_indices = nullptr, *d_data = nullptr;
    int rc = OFB_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_indptr, (dim + 1) * ib)) != cudaSuccess ||
        (e = cudaMalloc(&d_indices, std::max<size_t>(nnz, 1) * ib)) != cudaSuccess ||
        (e = cudaMalloc(&d_data, std::max<size_t>(nnz, 1) * 16)) != cudaSuccess) {
        rc = cuda_fail(e, "cudaMalloc(csc output)", __FILE__, __LINE__);
    }
    if (!rc) rc = ofb_csc_fill(p, d_indptr, d_indices, d_data, index_bits, 0);
    if (!rc) {
        if ((e = cudaMemcpy(h_indptr, d_indptr, (dim + 1) * ib, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (nnz && (e = cudaMemcpy(h_indices, d_indices, nnz * ib, cudaMemcpyDeviceToHost)) != cudaSuccess) ||
            (nnz && (e = cudaMemcpy(h_data, d_data, nnz * 16, cudaMemcpyDeviceToHost)) != cudaSuccess))
            rc = cuda_fail(e, "cudaMemcpy(csc output)", __FILE__, __LINE__);
    }
    cudaFree(d_indptr);
    cudaFree(d_indices);
    cudaFree(d_data);
    // scratch of this build is released; a new count is needed for another fill
    cudaFree(p->d_colptr);
    cudaFree(p->d_bitmap);
    p->d_colptr = nullptr;
    p->d_bitmap = nullptr;
    p->csc_nnz = -1;
    return rc;
}

}  // extern "C"
