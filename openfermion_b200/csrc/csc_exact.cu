// K4, scipy-order mode: per column, replay what scipy does to the triplets of
// qubit_operator_sparse / jordan_wigner_sparse (linalg/sparse_tools.py:126-133,187-193):
// the column's (row, value) pairs in term order are sorted by row with libstdc++'s std::sort
// (stdsort_emul.h replays its exact element movements), equal rows are summed left to
// right, exact zeros are dropped.  One thread owns one column; its T elements
// (row << 32 | term) live in a global scratch array interleaved across the columns of a
// batch, so that threads of a warp touch neighbouring words while they are in step.
#include <algorithm>
#include <cstdlib>

#include "ofb_internal.cuh"
#include "stdsort_emul.h"

namespace ofb {

constexpr int XB = 128;

// builds and sorts the column's element list; returns the element count
__device__ __forceinline__ int64_t build_sorted(const ScatterRow *__restrict__ rows, uint32_t T,
                                                uint64_t c, const ofb_sort::Seq &seq) {
    int64_t m = 0;
    for (uint32_t t = 0; t < T; ++t) {
        const ScatterRow r = rows[t];
        // a zero coefficient leaves no triplet in the reference (coo_matrix(0) is empty)
        if ((c & r.occ_mask) == r.occ_val && (r.cr != 0.0 || r.ci != 0.0)) {
            seq.set(m, ((c ^ r.x) << 32) | t);
            ++m;
        }
    }
    ofb_sort::std_sort(seq, m);
    return m;
}

__device__ __forceinline__ double2 row_value(const ScatterRow *__restrict__ rows, uint32_t t,
                                             uint64_t c, int n) {
    return scatter_value(rows[t], c, n);  // with the reference's signed zeros (ofb_internal.cuh)
}

// MODE 0: count non-zero sums; MODE 1: write them.  presorted != 0: the scratch still holds
// this column's sorted list and element count from the count pass.
template <int MODE, typename I>
__global__ void __launch_bounds__(XB)
k_csc_exact(const ScatterRow *__restrict__ rows, uint32_t T, uint64_t col0, uint64_t ncols,
            uint64_t *__restrict__ scratch, int64_t batch, int64_t *__restrict__ lens,
            int presorted, int64_t *__restrict__ counts, const int64_t *__restrict__ colptr,
            I *__restrict__ indices, double2 *__restrict__ data, int n) {
    const uint64_t b = (uint64_t)blockIdx.x * XB + threadIdx.x;
    if (b >= ncols) return;
    const uint64_t c = col0 + b;
    ofb_sort::Seq seq{scratch + b, batch};
    int64_t m;
    if (presorted) {
        m = lens[b];
    } else {
        m = build_sorted(rows, T, c, seq);
        lens[b] = m;
    }
    int64_t out = MODE == 1 ? colptr[c] : 0;
    int64_t i = 0;
    while (i < m) {
        const uint64_t e = seq.get(i);
        const uint32_t key = (uint32_t)(e >> 32);
        double2 acc = row_value(rows, (uint32_t)e, c, n);
        ++i;
        while (i < m) {
            const uint64_t f = seq.get(i);
            if ((uint32_t)(f >> 32) != key) break;
            const double2 v = row_value(rows, (uint32_t)f, c, n);
            acc.x += v.x;
            acc.y += v.y;
            ++i;
        }
        if (acc.x != 0.0 || acc.y != 0.0) {
            if (MODE == 1) {
                indices[out] = (I)key;
                data[out] = acc;
            }
            ++out;
        }
    }
    if (MODE == 0) counts[c] = out;
}

static int ensure_sort_scratch(ofb_plan *p) {
    const uint64_t dim = 1ull << p->n_qubits;
    const uint64_t T = (uint64_t)std::max<int64_t>(p->n_terms, 1);
    if (p->d_sort_scratch) return OFB_OK;
    size_t free_b = 0, total_b = 0;
    OFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget = std::min<uint64_t>((uint64_t)(free_b / 4), 16ull << 30);
    uint64_t batch = dim;
    while (batch > 1024 && batch * (T * 8 + 8) > budget) batch >>= 1;
    if (const char *forced = getenv("OFB_SORT_BATCH")) {  // tests: force the multi-batch path
        uint64_t want = strtoull(forced, nullptr, 10);
        while (want >= 128 && batch > want) batch >>= 1;
    }
    if (batch * (T * 8 + 8) > budget && batch * (T * 8 + 8) > free_b / 2)
        return fail(OFB_ERR_NOMEM, "not enough device memory for the scipy-order CSC mode");
    OFB_CUDA(cudaMalloc(&p->d_sort_scratch, batch * (T * 8 + 8)));
    p->sort_batch = (int64_t)batch;
    return OFB_OK;
}

int csc_exact_count(ofb_plan *p, int64_t *d_counts, cudaStream_t s) {
    int rc = ensure_sort_scratch(p);
    if (rc) return rc;
    const uint64_t dim = 1ull << p->n_qubits;
    const uint64_t batch = (uint64_t)p->sort_batch;
    int64_t *lens = (int64_t *)(p->d_sort_scratch + batch * (uint64_t)std::max<int64_t>(p->n_terms, 1));
    for (uint64_t col0 = 0; col0 < dim; col0 += batch) {
        const uint64_t ncols = std::min(batch, dim - col0);
        k_csc_exact<0, int32_t><<<(unsigned)((ncols + XB - 1) / XB), XB, 0, s>>>(
            p->d_rows_orig, (uint32_t)p->n_terms, col0, ncols, p->d_sort_scratch, (int64_t)batch, lens,
            0, d_counts, nullptr, nullptr, nullptr, p->n_qubits);
        OFB_LAUNCH_CHECK();
    }
    return OFB_OK;
}

int csc_exact_fill(ofb_plan *p, const int64_t *d_colptr, void *d_indices, void *d_data,
                   int index_bits, cudaStream_t s) {
    int rc = ensure_sort_scratch(p);
    if (rc) return rc;
    const uint64_t dim = 1ull << p->n_qubits;
    const uint64_t batch = (uint64_t)p->sort_batch;
    int64_t *lens = (int64_t *)(p->d_sort_scratch + batch * (uint64_t)std::max<int64_t>(p->n_terms, 1));
    const int presorted = batch >= dim ? 1 : 0;  // one batch: the count pass left everything sorted
    for (uint64_t col0 = 0; col0 < dim; col0 += batch) {
        const uint64_t ncols = std::min(batch, dim - col0);
        const unsigned grid = (unsigned)((ncols + XB - 1) / XB);
        if (index_bits == 32)
            k_csc_exact<1, int32_t><<<grid, XB, 0, s>>>(
                p->d_rows_orig, (uint32_t)p->n_terms, col0, ncols, p->d_sort_scratch, (int64_t)batch,
                lens, presorted, nullptr, d_colptr, (int32_t *)d_indices, (double2 *)d_data, p->n_qubits);
        else
            k_csc_exact<1, int64_t><<<grid, XB, 0, s>>>(
                p->d_rows_orig, (uint32_t)p->n_terms, col0, ncols, p->d_sort_scratch, (int64_t)batch,
                lens, presorted, nullptr, d_colptr, (int64_t *)d_indices, (double2 *)d_data, p->n_qubits);
        OFB_LAUNCH_CHECK();
    }
    return OFB_OK;
}

}  // namespace ofb
