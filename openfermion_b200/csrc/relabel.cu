// GF(2)-linear relabelling of basis indices on the device (openfermion_b200/shard_basis.py):
// c = B^-1 c' is the XOR of the columns of B^-1 selected by the set bits of c'.  Two users:
//   * ofb_vec_fill_random_mapped: a shard of the counter-based random vector keyed by the
//     NATURAL (reference-order) index, so that every sharding / basis / GPU count holds the
//     same state (bench.py cross-checks <psi|H|psi> and sampled outputs across N);
//   * ofb_vec_relabel: psi'[c'] = psi[B^-1 c'] and back, so that a one-GPU solve can run in a
//     locality-ordered basis and still take and return vectors in the reference's order
//     (qubit q = index bit n-1-q, linalg/linear_qubit_operator.py:78-84).
#include "ofb_internal.cuh"

namespace ofb {

constexpr int RB = 256;
struct ColTable {
    uint64_t col[40];
};

__device__ __forceinline__ uint64_t relabel_index(const ColTable &t, uint64_t c, int n_bits) {
    uint64_t out = 0;
#pragma unroll 8
    for (int b = 0; b < n_bits; ++b) out ^= ((c >> b) & 1ull) ? t.col[b] : 0ull;
    return out;
}

__device__ __forceinline__ uint64_t rl_splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// same generator as k_fill_random (vec.cu), keyed by the mapped index
__global__ void __launch_bounds__(RB) k_fill_random_mapped(int64_t n, uint64_t seed, int real_only,
                                                           double2 *__restrict__ x, uint64_t index_base,
                                                           ColTable table, int n_bits) {
    const int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = relabel_index(table, index_base | (uint64_t)i, n_bits);
    const uint64_t a = rl_splitmix64(seed ^ rl_splitmix64(2 * g));
    const uint64_t b = rl_splitmix64(seed ^ rl_splitmix64(2 * g + 1));
    const double scale = 1.0 / 9007199254740992.0;  // 2^-53
    x[i] = make_double2((double)(a >> 11) * scale, real_only ? 0.0 : (double)(b >> 11) * scale);
}

// out[c'] = in[map(c')]  (gather; `map` = columns of B^-1 gives psi' from psi, columns of B the way back)
__global__ void __launch_bounds__(RB) k_relabel(int64_t n, const double2 *__restrict__ in,
                                                double2 *__restrict__ out, ColTable table, int n_bits) {
    const int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x;
    if (i >= n) return;
    out[i] = in[relabel_index(table, (uint64_t)i, n_bits)];
}

static int load_table(ColTable *t, const uint64_t *h_cols, int n_bits) {
    if (!h_cols || n_bits < 0 || n_bits > 40) return fail(OFB_ERR_INVALID, "bad column table");
    for (int b = 0; b < 40; ++b) t->col[b] = b < n_bits ? h_cols[b] : 0ull;
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_vec_fill_random_mapped(int64_t n, uint64_t seed, int real_only, void *d_x, uint64_t index_base,
                               int n_bits, const uint64_t *h_cols, void *stream) {
    if (n < 0 || !d_x) return fail(OFB_ERR_INVALID, "bad argument");
    ColTable t;
    int rc = load_table(&t, h_cols, n_bits);
    if (rc) return rc;
    if (n == 0) return OFB_OK;
    k_fill_random_mapped<<<(unsigned)((n + RB - 1) / RB), RB, 0, (cudaStream_t)stream>>>(
        n, seed, real_only, (double2 *)d_x, index_base, t, n_bits);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_vec_relabel(int n_bits, const uint64_t *h_cols, const void *d_in, void *d_out, void *stream) {
    if (!d_in || !d_out || d_in == d_out) return fail(OFB_ERR_INVALID, "bad argument");
    ColTable t;
    int rc = load_table(&t, h_cols, n_bits);
    if (rc) return rc;
    const int64_t n = 1ll << n_bits;
    k_relabel<<<(unsigned)((n + RB - 1) / RB), RB, 0, (cudaStream_t)stream>>>(
        n, (const double2 *)d_in, (double2 *)d_out, t, n_bits);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

}  // extern "C"
