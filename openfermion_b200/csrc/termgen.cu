// N3: fermion-to-qubit term generation on the device.
//
// The reference's jordan_wigner / bravyi_kitaev multiply symbolic QubitOperators term by term in
// Python (transforms/opconversions/jordan_wigner.py:132-331, bravyi_kitaev.py:242-337; 4.2 s for
// the 88 495-string 28-qubit molecular Hamiltonian).  Under either encoding a ladder operator is
// a pair of Pauli strings,  a_j^(dagger) = 1/2 (A_j -/+ i B_j),  given per mode as (x, z) masks;
// a product of k ladder operators is the sum of its 2^k string products
//     P(x1,z1) P(x2,z2) = i^{|x1&z1| + |x2&z2| + 2|z1&x2| - |x3&z3|} P(x1^x2, z1^z2).
// One thread expands one (term, choice) pair; equal strings are then merged with a STABLE radix
// sort (CUB) and a sequential sum per run -- the same summation order as the host builder's
// numpy.bincount (openfermion_b200/hamiltonians.py:_collect), so both give the same coefficients
// bit for bit -- and sums below the reference's 1e-8 threshold are dropped (symbolic_operator.py
// += rule).  Output: strings sorted by (x, z), qubit-space masks (bit q = qubit q).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "ofb_internal.cuh"

struct ofb_termgen {
    int n_qubits = 0;
    int device = 0;
    uint64_t *d_tab = nullptr;  // xa, za, xb, zb: 4 * n_qubits
    // expanded entries in insertion order
    uint64_t *d_x = nullptr, *d_z = nullptr;
    double2 *d_c = nullptr;
    int64_t count = 0, capacity = 0;
    // merged result
    uint64_t *d_ox = nullptr, *d_oz = nullptr;
    double2 *d_oc = nullptr;
    int64_t n_out = -1;
};

namespace ofb {

constexpr int GB_ = 256;

__device__ __forceinline__ double2 cmul_exact(double2 a, double2 b) {
    // numpy's complex product, unfused: (ar br - ai bi, ar bi + ai br)
    return make_double2(__dadd_rn(__dmul_rn(a.x, b.x), -__dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}

__global__ void __launch_bounds__(GB_)
k_ladder_expand(int64_t n_terms, int k, uint32_t kinds, const int32_t *__restrict__ modes,
                const double2 *__restrict__ coeff, const uint64_t *__restrict__ tab, int n_qubits,
                uint64_t *__restrict__ out_x, uint64_t *__restrict__ out_z, double2 *__restrict__ out_c) {
    const int64_t idx = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (idx >= (n_terms << k)) return;
    const int64_t t = idx % n_terms;
    const uint32_t choice = (uint32_t)(idx / n_terms);  // factor 0 is the most significant bit
    const uint64_t *xa = tab, *za = tab + n_qubits, *xb = tab + 2 * n_qubits, *zb = tab + 3 * n_qubits;
    uint64_t x = 0, z = 0;
    double2 c = coeff[t];
    for (int j = 0; j < k; ++j) {
        const int mode = modes[t * k + j];
        const bool second = (choice >> (k - 1 - j)) & 1u;
        const bool raising = (kinds >> j) & 1u;
        const uint64_t xj = second ? xb[mode] : xa[mode], zj = second ? zb[mode] : za[mode];
        const double2 f = second ? make_double2(0.0, raising ? -0.5 : 0.5) : make_double2(0.5, 0.0);
        const uint64_t x3 = x ^ xj, z3 = z ^ zj;
        const int e = (__popcll(x & z) + __popcll(xj & zj) + 2 * __popcll(z & xj) - __popcll(x3 & z3)) & 3;
        const double2 phase = e == 0 ? make_double2(1.0, 0.0) : e == 1 ? make_double2(0.0, 1.0)
                              : e == 2 ? make_double2(-1.0, 0.0) : make_double2(0.0, -1.0);
        c = cmul_exact(cmul_exact(c, f), phase);
        x = x3;
        z = z3;
    }
    out_x[idx] = x;
    out_z[idx] = z;
    out_c[idx] = c;
}

__global__ void __launch_bounds__(GB_) k_iota(int64_t n, uint32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(GB_)
k_gather_keys(int64_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ src, int shift_hi,
              const uint64_t *__restrict__ src_lo, uint64_t *__restrict__ out) {
    // key of entry perm[i]: src alone, or (src << 32) | src_lo for the packed n <= 32 form
    const int64_t i = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = perm ? perm[i] : (uint32_t)i;
    out[i] = shift_hi ? ((src[p] << 32) | src_lo[p]) : src[p];
}
__global__ void __launch_bounds__(GB_)
k_heads(int64_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ x,
        const uint64_t *__restrict__ z, int32_t *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = perm[i];
    head[i] = (i == 0 || x[p] != x[perm[i - 1]] || z[p] != z[perm[i - 1]]) ? 1 : 0;
}
// one thread per run of equal strings: sequential sum in sorted (= insertion) order
__global__ void __launch_bounds__(GB_)
k_run_sums(int64_t n, const uint32_t *__restrict__ perm, const int32_t *__restrict__ head,
           const int32_t *__restrict__ slot, const uint64_t *__restrict__ x, const uint64_t *__restrict__ z,
           const double2 *__restrict__ c, double tol, uint64_t *__restrict__ rx, uint64_t *__restrict__ rz,
           double2 *__restrict__ rc, int32_t *__restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (i >= n || !head[i]) return;
    double re = 0.0, im = 0.0;
    int64_t j = i;
    do {
        const double2 v = c[perm[j]];
        re += v.x;
        im += v.y;
        ++j;
    } while (j < n && !head[j]);
    const int32_t s = slot[i];
    rx[s] = x[perm[i]];
    rz[s] = z[perm[i]];
    rc[s] = make_double2(re, im);
    keep[s] = hypot(re, im) >= tol ? 1 : 0;
}
__global__ void __launch_bounds__(GB_)
k_compact(int64_t n, const int32_t *__restrict__ keep, const int32_t *__restrict__ slot,
          const uint64_t *__restrict__ rx, const uint64_t *__restrict__ rz, const double2 *__restrict__ rc,
          uint64_t *__restrict__ ox, uint64_t *__restrict__ oz, double2 *__restrict__ oc) {
    const int64_t i = (int64_t)blockIdx.x * GB_ + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int32_t s = slot[i];
    ox[s] = rx[i];
    oz[s] = rz[i];
    oc[s] = rc[i];
}

static inline unsigned tg_grid(int64_t n) { return (unsigned)std::max<int64_t>(1, (n + GB_ - 1) / GB_); }

struct TgScratch {  // frees everything on scope exit
    std::vector<void *> ptrs;
    ~TgScratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    int alloc(T **p, size_t count) {
        *p = nullptr;
        OFB_CUDA(cudaMalloc(p, std::max<size_t>(count, 1) * sizeof(T)));
        ptrs.push_back(*p);
        return OFB_OK;
    }
};

static int tg_reserve(ofb_termgen *g, int64_t extra) {
    if (g->count + extra <= g->capacity) return OFB_OK;
    const int64_t cap = std::max<int64_t>(g->count + extra, 2 * g->capacity);
    uint64_t *nx = nullptr, *nz = nullptr;
    double2 *nc = nullptr;
    OFB_CUDA(cudaMalloc(&nx, cap * sizeof(uint64_t)));
    OFB_CUDA(cudaMalloc(&nz, cap * sizeof(uint64_t)));
    OFB_CUDA(cudaMalloc(&nc, cap * sizeof(double2)));
    if (g->count) {
        OFB_CUDA(cudaMemcpy(nx, g->d_x, g->count * sizeof(uint64_t), cudaMemcpyDeviceToDevice));
        OFB_CUDA(cudaMemcpy(nz, g->d_z, g->count * sizeof(uint64_t), cudaMemcpyDeviceToDevice));
        OFB_CUDA(cudaMemcpy(nc, g->d_c, g->count * sizeof(double2), cudaMemcpyDeviceToDevice));
    }
    cudaFree(g->d_x);
    cudaFree(g->d_z);
    cudaFree(g->d_c);
    g->d_x = nx;
    g->d_z = nz;
    g->d_c = nc;
    g->capacity = cap;
    return OFB_OK;
}

// stable sort of perm by key (cub radix sort of (key, perm) pairs)
static int tg_sort_by(int64_t n, uint64_t *keys, uint64_t *keys_alt, uint32_t **perm, uint32_t **perm_alt,
                      int end_bit, TgScratch &scratch) {
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, keys_alt, *perm, *perm_alt, (int)n, 0, end_bit);
    void *temp = nullptr;
    int rc = scratch.alloc((char **)&temp, temp_bytes);
    if (rc) return rc;
    OFB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_alt, *perm, *perm_alt, (int)n, 0, end_bit));
    g_launches.fetch_add(1);
    std::swap(*perm, *perm_alt);
    return OFB_OK;
}

static int tg_scan(int64_t n, const int32_t *in, int32_t *out, TgScratch &scratch) {
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, in, out, (int)n);
    void *temp = nullptr;
    int rc = scratch.alloc((char **)&temp, temp_bytes);
    if (rc) return rc;
    OFB_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n));
    g_launches.fetch_add(1);
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_termgen_destroy(ofb_termgen *g) {
    if (!g) return OFB_OK;
    cudaSetDevice(g->device);
    cudaFree(g->d_tab);
    cudaFree(g->d_x);
    cudaFree(g->d_z);
    cudaFree(g->d_c);
    cudaFree(g->d_ox);
    cudaFree(g->d_oz);
    cudaFree(g->d_oc);
    cudaGetLastError();
    delete g;
    return OFB_OK;
}

int ofb_termgen_create(int n_qubits, const uint64_t *xa, const uint64_t *za, const uint64_t *xb,
                       const uint64_t *zb, int device, ofb_termgen **out) {
    if (!out || n_qubits < 1 || n_qubits > 64 || !xa || !za || !xb || !zb) return fail(OFB_ERR_INVALID, "bad argument");
    *out = nullptr;
    OFB_CUDA(cudaSetDevice(device));
    ofb_termgen *g = new (std::nothrow) ofb_termgen();
    if (!g) return fail(OFB_ERR_NOMEM, "host allocation failed");
    g->n_qubits = n_qubits;
    g->device = device;
    cudaError_t e = cudaMalloc(&g->d_tab, 4 * (size_t)n_qubits * sizeof(uint64_t));
    const uint64_t *src[4] = {xa, za, xb, zb};
    for (int k = 0; k < 4 && e == cudaSuccess; ++k)
        e = cudaMemcpy(g->d_tab + (size_t)k * n_qubits, src[k], n_qubits * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ofb_termgen_destroy(g);
        return cuda_fail(e, "term generator tables", __FILE__, __LINE__);
    }
    *out = g;
    return OFB_OK;
}

int ofb_termgen_add(ofb_termgen *g, int k, uint32_t kinds_mask, int64_t n_terms, const int32_t *h_modes,
                    const double *h_coeff_ri) {
    if (!g || k < 0 || k > 16 || n_terms < 0 || (n_terms > 0 && !h_coeff_ri) || (n_terms > 0 && k > 0 && !h_modes))
        return fail(OFB_ERR_INVALID, "bad argument");
    if (n_terms == 0) return OFB_OK;
    if ((n_terms << k) + g->count >= (1ll << 31)) return fail(OFB_ERR_INVALID, "more than 2^31 expanded strings");
    for (int64_t i = 0; i < n_terms * k; ++i)
        if (h_modes[i] < 0 || h_modes[i] >= g->n_qubits) return fail(OFB_ERR_INVALID, "mode index out of range");
    OFB_CUDA(cudaSetDevice(g->device));
    OFB_TRY
    TgScratch scratch;
    int32_t *d_modes = nullptr;
    double2 *d_coeff = nullptr;
    int rc;
    if ((rc = scratch.alloc(&d_modes, (size_t)n_terms * std::max(k, 1))) || (rc = scratch.alloc(&d_coeff, (size_t)n_terms)))
        return rc;
    if (k > 0) OFB_CUDA(cudaMemcpy(d_modes, h_modes, (size_t)n_terms * k * sizeof(int32_t), cudaMemcpyHostToDevice));
    OFB_CUDA(cudaMemcpy(d_coeff, h_coeff_ri, (size_t)n_terms * sizeof(double2), cudaMemcpyHostToDevice));
    const int64_t extra = n_terms << k;
    if ((rc = tg_reserve(g, extra))) return rc;
    k_ladder_expand<<<tg_grid(extra), GB_>>>(n_terms, k, kinds_mask, d_modes, d_coeff, g->d_tab, g->n_qubits,
                                             g->d_x + g->count, g->d_z + g->count, g->d_c + g->count);
    OFB_LAUNCH_CHECK();
    OFB_CUDA(cudaDeviceSynchronize());
    g->count += extra;
    g->n_out = -1;
    return OFB_OK;
    OFB_CATCH((void)0)
}

int ofb_termgen_collect(ofb_termgen *g, double tol, int64_t *n_out) {
    if (!g || !n_out) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(g->device));
    OFB_TRY
    cudaFree(g->d_ox);
    cudaFree(g->d_oz);
    cudaFree(g->d_oc);
    g->d_ox = g->d_oz = nullptr;
    g->d_oc = nullptr;
    g->n_out = 0;
    *n_out = 0;
    const int64_t n = g->count;
    if (n == 0) return OFB_OK;
    TgScratch scratch;
    int rc;
    uint32_t *perm = nullptr, *perm_alt = nullptr;
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    int32_t *head = nullptr, *slot = nullptr, *keep = nullptr, *slot2 = nullptr;
    if ((rc = scratch.alloc(&perm, n)) || (rc = scratch.alloc(&perm_alt, n)) || (rc = scratch.alloc(&keys, n)) ||
        (rc = scratch.alloc(&keys_alt, n)) || (rc = scratch.alloc(&head, n + 1)) || (rc = scratch.alloc(&slot, n + 1)))
        return rc;
    k_iota<<<tg_grid(n), GB_>>>(n, perm);
    OFB_LAUNCH_CHECK();
    if (g->n_qubits <= 32) {  // one pass on the packed key (x << 32) | z
        k_gather_keys<<<tg_grid(n), GB_>>>(n, nullptr, g->d_x, 1, g->d_z, keys);
        OFB_LAUNCH_CHECK();
        if ((rc = tg_sort_by(n, keys, keys_alt, &perm, &perm_alt, 64, scratch))) return rc;
    } else {  // lexicographic (x, z): stable sort by z, then by x
        k_gather_keys<<<tg_grid(n), GB_>>>(n, nullptr, g->d_z, 0, nullptr, keys);
        OFB_LAUNCH_CHECK();
        if ((rc = tg_sort_by(n, keys, keys_alt, &perm, &perm_alt, 64, scratch))) return rc;
        k_gather_keys<<<tg_grid(n), GB_>>>(n, perm, g->d_x, 0, nullptr, keys);
        OFB_LAUNCH_CHECK();
        if ((rc = tg_sort_by(n, keys, keys_alt, &perm, &perm_alt, 64, scratch))) return rc;
    }
    k_heads<<<tg_grid(n), GB_>>>(n, perm, g->d_x, g->d_z, head);
    OFB_LAUNCH_CHECK();
    if ((rc = tg_scan(n, head, slot, scratch))) return rc;
    int32_t last_slot = 0, last_head = 0;
    OFB_CUDA(cudaMemcpy(&last_slot, slot + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    OFB_CUDA(cudaMemcpy(&last_head, head + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int64_t runs = (int64_t)last_slot + last_head;
    uint64_t *rx = nullptr, *rz = nullptr;
    double2 *rc_ = nullptr;
    if ((rc = scratch.alloc(&rx, runs)) || (rc = scratch.alloc(&rz, runs)) || (rc = scratch.alloc(&rc_, runs)) ||
        (rc = scratch.alloc(&keep, runs + 1)) || (rc = scratch.alloc(&slot2, runs + 1)))
        return rc;
    k_run_sums<<<tg_grid(n), GB_>>>(n, perm, head, slot, g->d_x, g->d_z, g->d_c, tol, rx, rz, rc_, keep);
    OFB_LAUNCH_CHECK();
    if ((rc = tg_scan(runs, keep, slot2, scratch))) return rc;
    int32_t s_last = 0, k_last = 0;
    OFB_CUDA(cudaMemcpy(&s_last, slot2 + (runs - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    OFB_CUDA(cudaMemcpy(&k_last, keep + (runs - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int64_t kept = (int64_t)s_last + k_last;
    OFB_CUDA(cudaMalloc(&g->d_ox, std::max<int64_t>(kept, 1) * sizeof(uint64_t)));
    OFB_CUDA(cudaMalloc(&g->d_oz, std::max<int64_t>(kept, 1) * sizeof(uint64_t)));
    OFB_CUDA(cudaMalloc(&g->d_oc, std::max<int64_t>(kept, 1) * sizeof(double2)));
    k_compact<<<tg_grid(runs), GB_>>>(runs, keep, slot2, rx, rz, rc_, g->d_ox, g->d_oz, g->d_oc);
    OFB_LAUNCH_CHECK();
    OFB_CUDA(cudaDeviceSynchronize());
    g->n_out = kept;
    *n_out = kept;
    return OFB_OK;
    OFB_CATCH((void)0)
}

int ofb_termgen_fetch(ofb_termgen *g, uint64_t *h_x, uint64_t *h_z, double *h_coeff_ri) {
    if (!g) return fail(OFB_ERR_INVALID, "NULL argument");
    if (g->n_out < 0) return fail(OFB_ERR_STATE, "ofb_termgen_collect must run first");
    if (g->n_out == 0) return OFB_OK;
    if (!h_x || !h_z || !h_coeff_ri) return fail(OFB_ERR_INVALID, "NULL output");
    OFB_CUDA(cudaSetDevice(g->device));
    OFB_CUDA(cudaMemcpy(h_x, g->d_ox, g->n_out * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    OFB_CUDA(cudaMemcpy(h_z, g->d_oz, g->n_out * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    OFB_CUDA(cudaMemcpy(h_coeff_ri, g->d_oc, g->n_out * sizeof(double2), cudaMemcpyDeviceToHost));
    return OFB_OK;
}

}  // extern "C"
