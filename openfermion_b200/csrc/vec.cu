// K5: complex128 vector kernels for the device-resident Krylov loops, with
// deterministic two-stage reductions (fixed grid, fixed summation order).
#include <cmath>
#include <mutex>

#include "ofb_internal.cuh"

namespace ofb {

constexpr int VB = 256;        // threads per block
constexpr int MAX_RBLOCKS = 148 * 8;  // reduction grid (multiple of the SM count)
constexpr int MD = 8;          // columns per multi-dot / multi-axpy pass

struct DevScratch {
    double *d = nullptr;
    size_t bytes = 0;
    double *pinned = nullptr;
};
static DevScratch g_scratch[64];
static std::mutex g_scratch_mu;
// The reductions share one scratch buffer and one pinned result buffer per device.  Every entry
// point that uses them holds this lock from its first kernel to the stream synchronisation that
// ends it, so calls from several host threads or streams on one device serialise instead of
// racing on the partial sums.
static std::mutex g_reduce_mu[64];
struct ReduceLock {
    std::unique_lock<std::mutex> lk;
    ReduceLock() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        lk = std::unique_lock<std::mutex>(g_reduce_mu[dev]);
    }
};

double *global_scratch(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    DevScratch &s = g_scratch[dev];
    if (s.bytes < bytes) {
        if (s.d) cudaFree(s.d);
        s.d = nullptr;
        s.bytes = 0;
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        if (cudaMalloc(&s.d, want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        s.bytes = want;
    }
    return s.d;
}

double *global_pinned() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    DevScratch &s = g_scratch[dev];
    if (!s.pinned) {
        if (cudaMallocHost(&s.pinned, 4096 * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    return s.pinned;
}

static inline unsigned rgrid(int64_t n) {
    int64_t b = (n + VB - 1) / VB;
    if (b < 1) b = 1;
    return (unsigned)(b < MAX_RBLOCKS ? b : MAX_RBLOCKS);
}
static inline unsigned egrid(int64_t n, int per_thread = 1) {
    int64_t b = (n + (int64_t)VB * per_thread - 1) / ((int64_t)VB * per_thread);
    return (unsigned)(b < 1 ? 1 : b);
}

// block-wide fixed-order reduction of W values per thread; result valid in thread 0
template <int W, bool MAX>
__device__ __forceinline__ void block_reduce(double (&v)[W], double *sm /* W * VB/32 */) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double o = __shfl_xor_sync(0xffffffffu, v[w], off);
            v[w] = MAX ? fmax(v[w], o) : v[w] + o;
        }
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int w = 0; w < W; ++w) sm[warp * W + w] = v[w];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            double t = sm[w];
            for (int k = 1; k < VB / 32; ++k) t = MAX ? fmax(t, sm[k * W + w]) : t + sm[k * W + w];
            v[w] = t;
        }
    }
}

template <bool MAX>
__global__ void __launch_bounds__(VB) k_reduce_rows(const double *__restrict__ part, int64_t count,
                                                    int width, double *__restrict__ out) {
    __shared__ double sm[VB / 32];
    for (int w = 0; w < width; ++w) {
        double v[1] = {MAX ? 0.0 : 0.0};
        for (int64_t i = threadIdx.x; i < count; i += VB) {
            double x = part[i * width + w];
            v[0] = MAX ? fmax(v[0], x) : v[0] + x;
        }
        block_reduce<1, MAX>(v, sm);
        if (threadIdx.x == 0) out[w] = v[0];
        __syncthreads();
    }
}

int reduce_partials(const double *d_partials, int64_t count, int width, double *d_out,
                    cudaStream_t s) {
    k_reduce_rows<false><<<1, VB, 0, s>>>(d_partials, count, width, d_out);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
static int reduce_partials_max(const double *d_partials, int64_t count, int width, double *d_out,
                               cudaStream_t s) {
    k_reduce_rows<true><<<1, VB, 0, s>>>(d_partials, count, width, d_out);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

__global__ void __launch_bounds__(VB) k_dotc(const double2 *__restrict__ x,
                                             const double2 *__restrict__ y, int64_t n,
                                             double *__restrict__ part) {
    __shared__ double sm[2 * VB / 32];
    double v[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        double2 a = x[i], b = y[i];
        v[0] += a.x * b.x + a.y * b.y;
        v[1] += a.x * b.y - a.y * b.x;
    }
    block_reduce<2, false>(v, sm);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = v[0];
        part[2 * blockIdx.x + 1] = v[1];
    }
}

__global__ void __launch_bounds__(VB) k_sumsq(const double2 *__restrict__ x, int64_t n,
                                              double *__restrict__ part) {
    __shared__ double sm[VB / 32];
    double v[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        double2 a = x[i];
        v[0] += a.x * a.x + a.y * a.y;
    }
    block_reduce<1, false>(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(VB) k_maxabs(const double2 *__restrict__ x, int64_t n,
                                               double *__restrict__ part) {
    __shared__ double sm[VB / 32];
    double v[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        double2 a = x[i];
        v[0] = fmax(v[0], hypot(a.x, a.y));
    }
    block_reduce<1, true>(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(VB) k_axpy(int64_t n, double ar, double ai,
                                             const double2 *__restrict__ x, double2 *__restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    double2 a = x[i], b = y[i];
    b.x += ar * a.x - ai * a.y;
    b.y += ar * a.y + ai * a.x;
    y[i] = b;
}

__global__ void __launch_bounds__(VB) k_scal(int64_t n, double ar, double ai, double2 *__restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    double2 a = x[i];
    x[i] = make_double2(ar * a.x - ai * a.y, ar * a.y + ai * a.x);
}

struct Coef8 {
    double re[MD], im[MD];
};

__global__ void __launch_bounds__(VB) k_multi_dotc(int64_t n, int nc, const double2 *__restrict__ V,
                                                   int64_t ld, const double2 *__restrict__ w,
                                                   double *__restrict__ part) {
    __shared__ double sm[2 * MD * VB / 32];
    double v[2 * MD];
#pragma unroll
    for (int k = 0; k < 2 * MD; ++k) v[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        double2 b = w[i];
#pragma unroll
        for (int k = 0; k < MD; ++k) {
            if (k < nc) {
                double2 a = V[(int64_t)k * ld + i];
                v[2 * k] += a.x * b.x + a.y * b.y;
                v[2 * k + 1] += a.x * b.y - a.y * b.x;
            }
        }
    }
    block_reduce<2 * MD, false>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2 * MD; ++k) part[(int64_t)blockIdx.x * 2 * MD + k] = v[k];
    }
}

__global__ void __launch_bounds__(VB) k_multi_axpy(int64_t n, int nc, Coef8 a, const double2 *__restrict__ V,
                                                   int64_t ld, double br, double bi,
                                                   double2 *__restrict__ w) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    double2 o = w[i];
    double rr = br * o.x - bi * o.y, ri = br * o.y + bi * o.x;
    if (br == 0.0 && bi == 0.0) rr = ri = 0.0;  // beta = 0 overwrites (w may be uninitialised)
#pragma unroll
    for (int k = 0; k < MD; ++k) {
        if (k < nc) {
            double2 v = V[(int64_t)k * ld + i];
            rr += a.re[k] * v.x - a.im[k] * v.y;
            ri += a.re[k] * v.y + a.im[k] * v.x;
        }
    }
    w[i] = make_double2(rr, ri);
}

__global__ void __launch_bounds__(VB) k_lanczos_update(int64_t n, double alpha, double beta,
                                                       const double2 *__restrict__ v,
                                                       const double2 *__restrict__ vp,
                                                       double2 *__restrict__ w) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    double2 a = v[i], o = w[i];
    o.x -= alpha * a.x;
    o.y -= alpha * a.y;
    if (vp) {
        double2 b = vp[i];
        o.x -= beta * b.x;
        o.y -= beta * b.y;
    }
    w[i] = o;
}

__device__ __forceinline__ double dav_dinv(double d, double lambda, double eps) {
    double diff = d - lambda;
    return fabs(diff) > eps ? 1.0 / diff : 1.0 / eps;
}

// partial sums of v^H D r (2) and v^H D v (2)
__global__ void __launch_bounds__(VB) k_dav_dots(int64_t n, const double *__restrict__ diag,
                                                 double lambda, double eps,
                                                 const double2 *__restrict__ r,
                                                 const double2 *__restrict__ v,
                                                 double *__restrict__ part) {
    __shared__ double sm[4 * VB / 32];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        double di = dav_dinv(diag[i], lambda, eps);
        double2 a = v[i], b = r[i];
        double bx = di * b.x, by = di * b.y, ax = di * a.x, ay = di * a.y;
        acc[0] += a.x * bx + a.y * by;
        acc[1] += a.x * by - a.y * bx;
        acc[2] += a.x * ax + a.y * ay;
        acc[3] += a.x * ay - a.y * ax;
    }
    block_reduce<4, false>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) part[4 * blockIdx.x + k] = acc[k];
    }
}

__global__ void __launch_bounds__(VB) k_dav_newdir(int64_t n, const double *__restrict__ dots,
                                                   const double2 *__restrict__ r,
                                                   const double2 *__restrict__ v,
                                                   double2 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    // q = (v^H D r) / (v^H D v), complex division as numpy does it (Smith's algorithm
    // is not needed for parity at 1e-12: plain formula on well-scaled values)
    double nr = dots[0], ni = dots[1], dr = dots[2], dim_ = dots[3];
    double den = dr * dr + dim_ * dim_;
    double qr = (nr * dr + ni * dim_) / den, qi = (ni * dr - nr * dim_) / den;
    double2 a = v[i], b = r[i];
    out[i] = make_double2(-b.x + (a.x * qr - a.y * qi), -b.y + (a.x * qi + a.y * qr));
}

__global__ void __launch_bounds__(VB) k_drop_imag(int64_t n, double2 *__restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i < n) x[i].y = 0.0;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(VB) k_fill_random(int64_t n, uint64_t seed, int real_only,
                                                    double2 *__restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x;
    if (i >= n) return;
    uint64_t a = splitmix64(seed ^ splitmix64(2 * (uint64_t)i));
    uint64_t b = splitmix64(seed ^ splitmix64(2 * (uint64_t)i + 1));
    const double scale = 1.0 / 9007199254740992.0;  // 2^-53
    x[i] = make_double2((double)(a >> 11) * scale, real_only ? 0.0 : (double)(b >> 11) * scale);
}

// CSR y = A x, one warp per row (rows of Pauli-sum matrices hold tens to thousands of entries)
template <typename I>
__global__ void __launch_bounds__(VB) k_csr_matvec(int64_t n_rows, const I *__restrict__ indptr,
                                                   const I *__restrict__ indices,
                                                   const double2 *__restrict__ data,
                                                   const double2 *__restrict__ x,
                                                   double2 *__restrict__ y) {
    int64_t row = ((int64_t)blockIdx.x * VB + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    double sr = 0.0, si = 0.0;
    for (int64_t k = (int64_t)indptr[row] + lane; k < (int64_t)indptr[row + 1]; k += 32) {
        double2 a = data[k], b = x[(int64_t)indices[k]];
        sr += a.x * b.x - a.y * b.y;
        si += a.x * b.y + a.y * b.x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, off);
        si += __shfl_xor_sync(0xffffffffu, si, off);
    }
    if (lane == 0) y[row] = make_double2(sr, si);
}

static int fetch_scalars(const double *d_src, int count, double *h_out, cudaStream_t s) {
    double *pin = global_pinned();
    if (!pin) return fail(OFB_ERR_NOMEM, "pinned scratch allocation failed");
    OFB_CUDA(cudaMemcpyAsync(pin, d_src, count * sizeof(double), cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < count; ++k) h_out[k] = pin[k];
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_vec_dotc(int64_t n, const void *d_x, const void *d_y, double *h_out_ri, void *stream) {
    if (n < 0 || !h_out_ri) return fail(OFB_ERR_INVALID, "bad argument");
    ReduceLock guard;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = rgrid(n);
    double *part = global_scratch((size_t)(2 * g + 2) * sizeof(double));
    if (!part) return fail(OFB_ERR_NOMEM, "scratch allocation failed");
    k_dotc<<<g, VB, 0, s>>>((const double2 *)d_x, (const double2 *)d_y, n, part);
    OFB_LAUNCH_CHECK();
    int rc = reduce_partials(part, g, 2, part + 2 * g, s);
    if (rc) return rc;
    return fetch_scalars(part + 2 * g, 2, h_out_ri, s);
}

int ofb_vec_nrm2(int64_t n, const void *d_x, double *h_out, void *stream) {
    if (n < 0 || !h_out) return fail(OFB_ERR_INVALID, "bad argument");
    ReduceLock guard;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = rgrid(n);
    double *part = global_scratch((size_t)(g + 1) * sizeof(double));
    if (!part) return fail(OFB_ERR_NOMEM, "scratch allocation failed");
    k_sumsq<<<g, VB, 0, s>>>((const double2 *)d_x, n, part);
    OFB_LAUNCH_CHECK();
    int rc = reduce_partials(part, g, 1, part + g, s);
    if (rc) return rc;
    rc = fetch_scalars(part + g, 1, h_out, s);
    if (rc) return rc;
    *h_out = std::sqrt(*h_out);
    return OFB_OK;
}

int ofb_vec_maxabs(int64_t n, const void *d_x, double *h_out, void *stream) {
    if (n < 0 || !h_out) return fail(OFB_ERR_INVALID, "bad argument");
    ReduceLock guard;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = rgrid(n);
    double *part = global_scratch((size_t)(g + 1) * sizeof(double));
    if (!part) return fail(OFB_ERR_NOMEM, "scratch allocation failed");
    k_maxabs<<<g, VB, 0, s>>>((const double2 *)d_x, n, part);
    OFB_LAUNCH_CHECK();
    int rc = reduce_partials_max(part, g, 1, part + g, s);
    if (rc) return rc;
    return fetch_scalars(part + g, 1, h_out, s);
}

int ofb_vec_axpy(int64_t n, const double *alpha_ri, const void *d_x, void *d_y, void *stream) {
    if (n < 0 || !alpha_ri) return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    k_axpy<<<egrid(n), VB, 0, (cudaStream_t)stream>>>(n, alpha_ri[0], alpha_ri[1],
                                                     (const double2 *)d_x, (double2 *)d_y);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_vec_scal(int64_t n, const double *alpha_ri, void *d_x, void *stream) {
    if (n < 0 || !alpha_ri) return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    k_scal<<<egrid(n), VB, 0, (cudaStream_t)stream>>>(n, alpha_ri[0], alpha_ri[1], (double2 *)d_x);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_vec_multi_dotc(int64_t n, int ncols, const void *d_V, int64_t ld, const void *d_w,
                       double *h_out_ri, void *stream) {
    if (n < 0 || ncols < 0 || !h_out_ri) return fail(OFB_ERR_INVALID, "bad argument");
    ReduceLock guard;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = rgrid(n);
    double *part = global_scratch((size_t)(2 * MD) * (g + 1) * sizeof(double));
    if (!part) return fail(OFB_ERR_NOMEM, "scratch allocation failed");
    for (int c0 = 0; c0 < ncols; c0 += MD) {
        int nc = ncols - c0 < MD ? ncols - c0 : MD;
        k_multi_dotc<<<g, VB, 0, s>>>(n, nc, (const double2 *)d_V + (int64_t)c0 * ld, ld,
                                      (const double2 *)d_w, part);
        OFB_LAUNCH_CHECK();
        int rc = reduce_partials(part, g, 2 * MD, part + (size_t)2 * MD * g, s);
        if (rc) return rc;
        double tmp[2 * MD];
        rc = fetch_scalars(part + (size_t)2 * MD * g, 2 * MD, tmp, s);
        if (rc) return rc;
        for (int k = 0; k < 2 * nc; ++k) h_out_ri[2 * c0 + k] = tmp[k];
    }
    return OFB_OK;
}

int ofb_vec_multi_axpy(int64_t n, int ncols, const double *a_ri, const void *d_V, int64_t ld,
                       const double *beta_ri, void *d_w, void *stream) {
    if (n < 0 || ncols < 0 || (ncols > 0 && !a_ri) || !beta_ri)
        return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    double br = beta_ri[0], bi = beta_ri[1];
    int c0 = 0;
    do {
        int nc = ncols - c0 < MD ? ncols - c0 : MD;
        Coef8 a;
        for (int k = 0; k < MD; ++k) {
            a.re[k] = k < nc ? a_ri[2 * (c0 + k)] : 0.0;
            a.im[k] = k < nc ? a_ri[2 * (c0 + k) + 1] : 0.0;
        }
        k_multi_axpy<<<egrid(n), VB, 0, s>>>(n, nc, a, (const double2 *)d_V + (int64_t)c0 * ld, ld, br,
                                             bi, (double2 *)d_w);
        OFB_LAUNCH_CHECK();
        br = 1.0;
        bi = 0.0;
        c0 += MD;
    } while (c0 < ncols);
    return OFB_OK;
}

int ofb_vec_lanczos_update(int64_t n, double alpha, double beta, const void *d_v,
                           const void *d_vprev, void *d_w, void *stream) {
    if (n < 0) return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    k_lanczos_update<<<egrid(n), VB, 0, (cudaStream_t)stream>>>(
        n, alpha, beta, (const double2 *)d_v, (const double2 *)d_vprev, (double2 *)d_w);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_davidson_correction(int64_t n, const double *d_diag, double lambda, double eps,
                            const void *d_r, const void *d_v, void *d_new, void *stream) {
    if (n <= 0 || !d_diag) return fail(OFB_ERR_INVALID, "bad argument");
    ReduceLock guard;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = rgrid(n);
    double *part = global_scratch((size_t)4 * (g + 1) * sizeof(double));
    if (!part) return fail(OFB_ERR_NOMEM, "scratch allocation failed");
    k_dav_dots<<<g, VB, 0, s>>>(n, d_diag, lambda, eps, (const double2 *)d_r, (const double2 *)d_v, part);
    OFB_LAUNCH_CHECK();
    int rc = reduce_partials(part, g, 4, part + (size_t)4 * g, s);
    if (rc) return rc;
    k_dav_newdir<<<egrid(n), VB, 0, s>>>(n, part + (size_t)4 * g, (const double2 *)d_r,
                                         (const double2 *)d_v, (double2 *)d_new);
    OFB_LAUNCH_CHECK();
    OFB_CUDA(cudaStreamSynchronize(s));  // the scratch is free again when the lock is released
    return OFB_OK;
}

int ofb_vec_drop_imag(int64_t n, void *d_x, void *stream) {
    if (n < 0) return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    k_drop_imag<<<egrid(n), VB, 0, (cudaStream_t)stream>>>(n, (double2 *)d_x);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_vec_fill_random(int64_t n, uint64_t seed, int real_only, void *d_x, void *stream) {
    if (n < 0) return fail(OFB_ERR_INVALID, "bad argument");
    if (n == 0) return OFB_OK;
    k_fill_random<<<egrid(n), VB, 0, (cudaStream_t)stream>>>(n, seed, real_only, (double2 *)d_x);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

int ofb_csr_matvec(int64_t n_rows, const void *d_indptr, const void *d_indices, const void *d_data,
                   int index_bits, const void *d_x, void *d_y, void *stream) {
    if (n_rows < 0 || (index_bits != 32 && index_bits != 64))
        return fail(OFB_ERR_INVALID, "bad argument");
    if (n_rows == 0) return OFB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = egrid(n_rows * 32);
    if (index_bits == 32)
        k_csr_matvec<int32_t><<<g, VB, 0, s>>>(n_rows, (const int32_t *)d_indptr,
                                               (const int32_t *)d_indices, (const double2 *)d_data,
                                               (const double2 *)d_x, (double2 *)d_y);
    else
        k_csr_matvec<int64_t><<<g, VB, 0, s>>>(n_rows, (const int64_t *)d_indptr,
                                               (const int64_t *)d_indices, (const double2 *)d_data,
                                               (const double2 *)d_x, (double2 *)d_y);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

}  // extern "C"
