// K8: device-resident Lanczos behind the C ABI (get_ground_state / get_gap,
// linalg/sparse_tools.py:566-587,1078-1095 -- the reference hands the operator to ARPACK).
//
// The recurrence runs on UNNORMALISED vectors u_k = beta_k v_k, so that no pass over a vector
// is spent on scaling and every scalar stays on the device between kernels:
//     w      = H u_k                                   (operator kernel; for a plan operator the
//     d_k    = <u_k, w>                                 same launch also leaves the partial sums
//     alpha_k = d_k / beta_k^2                          of this dot product: MODE 2 of K1)
//     u_{k+1} = w / beta_k - (alpha_k / beta_k) u_k - (beta_k / beta_{k-1}) u_{k-1}
//     beta_{k+1} = |u_{k+1}|,   [ y += (s_k / beta_k) u_k   when the Ritz vector is rebuilt ]
// One fused pass (k_lanczos_step: 3 vectors read, 1 written, norm partials) replaces the five
// vector kernels and two host round trips of a textbook iteration.  alpha_k / beta_k are kept in
// device history arrays and fetched only when the host checks convergence.
//
// Two layers:
//   * ofb_kws_*: the workspace + per-iteration primitives.  The sharded solver
//     (openfermion_b200/distributed.py) drives them from Python and all-reduces the two scalar
//     slots between kernels (NCCL, stream-ordered, no host sync);
//   * ofb_lanczos_lowest: the complete restarted solver for one device (plan, sector or CSR
//     operator), Ritz problem of the tridiagonal matrix by bisection + inverse iteration.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ofb_internal.cuh"

struct ofb_linop {
    int kind = 0;  // 0 plan (K1), 1 sector (K7), 2 CSR
    int device = 0;
    int64_t dim = 0;
    ofb_plan *plan = nullptr;
    ofb_sector *sector = nullptr;
    const void *indptr = nullptr, *indices = nullptr, *data = nullptr;
    int index_bits = 32;
};

struct ofb_kws {
    int device = 0;
    int64_t n = 0;
    int max_steps = 0;
    double *d_scal = nullptr;   // scalar slots, see enum below
    double *d_alpha = nullptr;  // alpha_k, k < max_steps
    double *d_beta = nullptr;   // beta_k, k <= max_steps (beta_0 = |start|)
    double *d_part = nullptr;
    size_t part_doubles = 0;
    double *h_pin = nullptr;  // 2 * (max_steps + 1) + SLOTS doubles
};

namespace ofb {

enum Slot { S_DOT_RE = 0, S_DOT_IM = 1, S_NRM2 = 2, S_BETA = 3, S_BETA_PREV = 4, S_C0 = 5, S_C1 = 6,
            S_C2 = 7, S_C3 = 8, S_FLAG = 9, S_SCALE = 10, SLOTS = 16 };

constexpr int KB = 256;
constexpr int KGRID = 148 * 8;

// coefficients of the fused update for step k.  replay: alpha / beta come from the history
// written by the first pass (bit-identical recurrence, no reductions needed).
__global__ void k_kws_alpha(double *__restrict__ s, double *__restrict__ alphas,
                            const double *__restrict__ betas, int k, int replay, double ritz_k) {
    if (threadIdx.x | blockIdx.x) return;
    double beta, beta_prev, alpha;
    if (replay) {
        beta = betas[k];
        beta_prev = k ? betas[k - 1] : 0.0;
        alpha = alphas[k];
    } else {
        beta = s[S_BETA];
        beta_prev = s[S_BETA_PREV];
        alpha = s[S_DOT_RE] / (beta * beta);
        alphas[k] = alpha;
        s[S_SCALE] = fmax(s[S_SCALE], fabs(alpha));
    }
    const bool dead = !(beta > 0.0) || (!replay && s[S_FLAG] != 0.0);
    s[S_C0] = dead ? 0.0 : 1.0 / beta;
    s[S_C1] = dead ? 0.0 : -alpha / beta;
    s[S_C2] = (dead || k == 0) ? 0.0 : -beta / beta_prev;
    s[S_C3] = dead ? 0.0 : ritz_k / beta;
}

// beta_{k+1} from the reduced norm; an (almost) invariant subspace raises the flag once and
// freezes the recurrence (all later coefficients are zero, no NaN is ever produced)
__global__ void k_kws_beta(double *__restrict__ s, double *__restrict__ betas, int k) {
    if (threadIdx.x | blockIdx.x) return;
    const double b = sqrt(fmax(s[S_NRM2], 0.0));
    betas[k + 1] = b;
    s[S_BETA_PREV] = s[S_BETA];
    s[S_BETA] = b;
    if (s[S_FLAG] == 0.0 && !(b > 1e-14 * fmax(1.0, s[S_SCALE]))) s[S_FLAG] = (double)(k + 1);
}

template <bool NORM, bool RITZ, bool PREV>
__global__ void __launch_bounds__(KB) k_lanczos_step(int64_t n, double2 *__restrict__ w,
                                                     const double2 *__restrict__ u,
                                                     const double2 *__restrict__ up,
                                                     double2 *__restrict__ y,
                                                     const double *__restrict__ s,
                                                     double *__restrict__ part) {
    const double c0 = s[S_C0], c1 = s[S_C1], c2 = s[S_C2], c3 = s[S_C3];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
        const double2 a = w[i], b = u[i];
        double rx = c0 * a.x + c1 * b.x, ry = c0 * a.y + c1 * b.y;
        if (PREV) {
            const double2 c = up[i];
            rx += c2 * c.x;
            ry += c2 * c.y;
        }
        w[i] = make_double2(rx, ry);
        if (RITZ) {
            double2 t = y[i];
            t.x += c3 * b.x;
            t.y += c3 * b.y;
            y[i] = t;
        }
        if (NORM) acc += rx * rx + ry * ry;
    }
    if (NORM) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        __shared__ double sm[KB / 32];
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < KB / 32; ++k) t += sm[k];
            part[blockIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(KB) k_kws_dotc(const double2 *__restrict__ x,
                                                 const double2 *__restrict__ y, int64_t n,
                                                 double *__restrict__ part) {
    double re = 0.0, im = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
        const double2 a = x[i], b = y[i];
        re += a.x * b.x + a.y * b.y;
        im += a.x * b.y - a.y * b.x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, off);
        im += __shfl_xor_sync(0xffffffffu, im, off);
    }
    __shared__ double sm[2 * KB / 32];
    if ((threadIdx.x & 31) == 0) {
        sm[2 * (threadIdx.x >> 5)] = re;
        sm[2 * (threadIdx.x >> 5) + 1] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int k = 0; k < KB / 32; ++k) {
            tr += sm[2 * k];
            ti += sm[2 * k + 1];
        }
        part[2 * blockIdx.x] = tr;
        part[2 * blockIdx.x + 1] = ti;
    }
}

// y = alpha * x (real alpha read from a device slot is not needed here: host value)
__global__ void __launch_bounds__(KB) k_scale_copy(int64_t n, double a, const double2 *__restrict__ x,
                                                   double2 *__restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x;
    if (i < n) {
        double2 v = x[i];
        y[i] = make_double2(a * v.x, a * v.y);
    }
}

static inline unsigned kgrid(int64_t n) {
    int64_t b = (n + KB - 1) / KB;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, KGRID));
}

static int kws_ensure_part(ofb_kws *ws, size_t doubles) {
    if (ws->part_doubles >= doubles) return OFB_OK;
    if (ws->d_part) cudaFree(ws->d_part);
    ws->d_part = nullptr;
    ws->part_doubles = 0;
    OFB_CUDA(cudaMalloc(&ws->d_part, doubles * sizeof(double)));
    ws->part_doubles = doubles;
    return OFB_OK;
}

// ---- lowest eigenpair of a real symmetric tridiagonal matrix ---------------------------------
// (what scipy.linalg.eigh_tridiagonal(select='i', select_range=(0, 0)) computes with LAPACK's
// stebz + stein: bisection on the Sturm count, then inverse iteration.)
static int sturm_count(const std::vector<double> &d, const std::vector<double> &e2, double x,
                       double pivmin) {
    int count = 0;
    double q = d[0] - x;
    if (std::fabs(q) < pivmin) q = -pivmin;
    if (q < 0) ++count;
    for (size_t i = 1; i < d.size(); ++i) {
        q = d[i] - x - e2[i - 1] / q;
        if (std::fabs(q) < pivmin) q = -pivmin;
        if (q < 0) ++count;
    }
    return count;
}

void tridiag_lowest(const std::vector<double> &d, const std::vector<double> &e, double *value,
                    std::vector<double> *vector) {
    const size_t m = d.size();
    if (m == 1) {
        *value = d[0];
        vector->assign(1, 1.0);
        return;
    }
    std::vector<double> e2(m - 1);
    double lo = d[0], hi = d[0], emax = 0.0;
    for (size_t i = 0; i < m; ++i) {
        const double r = (i ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(e[i]) : 0.0);
        lo = std::min(lo, d[i] - r);
        hi = std::max(hi, d[i] + r);
    }
    for (size_t i = 0; i + 1 < m; ++i) {
        e2[i] = e[i] * e[i];
        emax = std::max(emax, e2[i]);
    }
    const double norm = std::max(std::fabs(lo), std::fabs(hi));
    const double pivmin = std::max(2.2250738585072014e-308 * std::max(1.0, emax), 1e-300);
    lo -= 2.0 * norm * 2.22e-16 * m + 2.0 * pivmin;
    hi += 2.0 * norm * 2.22e-16 * m + 2.0 * pivmin;
    // bisection for the smallest eigenvalue: largest x with count(x) == 0
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (sturm_count(d, e2, mid, pivmin) >= 1) hi = mid; else lo = mid;
    }
    const double lambda = 0.5 * (lo + hi);
    *value = lambda;
    // inverse iteration: (T - lambda I) x_{j+1} = x_j with partial pivoting
    const double eps = 2.220446049250313e-16;
    const double tiny = eps * std::max(norm, 1e-300);
    std::vector<double> x(m), a(m), b(m), f(m), sub(m), rhs(m);  // U: diagonal, first / second super-diagonal
    uint64_t state = 0x9e3779b97f4a7c15ull;
    for (size_t i = 0; i < m; ++i) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        x[i] = 0.5 + (double)(state >> 11) / 9007199254740992.0;
    }
    for (int iter = 0; iter < 3; ++iter) {
        // LU of the shifted matrix with row interchanges (dlagtf-style)
        for (size_t i = 0; i < m; ++i) {
            a[i] = d[i] - lambda;
            b[i] = i + 1 < m ? e[i] : 0.0;  // super-diagonal
            f[i] = 0.0;
        }
        for (size_t i = 0; i < m; ++i) sub[i] = i + 1 < m ? e[i] : 0.0;
        rhs = x;
        for (size_t i = 0; i + 1 < m; ++i) {
            if (std::fabs(a[i]) >= std::fabs(sub[i])) {
                if (std::fabs(a[i]) < tiny) a[i] = a[i] < 0 ? -tiny : tiny;
                const double mult = sub[i] / a[i];
                a[i + 1] -= mult * b[i];
                rhs[i + 1] -= mult * rhs[i];
                // f[i] stays 0
            } else {
                const double mult = a[i] / sub[i];
                // swap rows i and i+1
                const double ai = sub[i], bi = a[i + 1], fi = (i + 2 < m) ? b[i + 1] : 0.0;
                const double old_b = b[i];
                a[i] = ai;
                b[i] = bi;
                f[i] = fi;
                a[i + 1] = old_b - mult * bi;
                if (i + 2 < m) b[i + 1] = -mult * fi;
                const double r0 = rhs[i + 1], r1 = rhs[i];
                rhs[i] = r0;
                rhs[i + 1] = r1 - mult * r0;
            }
        }
        if (std::fabs(a[m - 1]) < tiny) a[m - 1] = a[m - 1] < 0 ? -tiny : tiny;
        // back substitution
        for (size_t ii = m; ii-- > 0;) {
            double t = rhs[ii];
            if (ii + 1 < m) t -= b[ii] * x[ii + 1];
            if (ii + 2 < m) t -= f[ii] * x[ii + 2];
            x[ii] = t / a[ii];
        }
        double nrm = 0.0, big = 0.0;
        for (size_t i = 0; i < m; ++i) big = std::max(big, std::fabs(x[i]));
        for (size_t i = 0; i < m; ++i) {
            x[i] /= big;
            nrm += x[i] * x[i];
        }
        nrm = std::sqrt(nrm);
        for (size_t i = 0; i < m; ++i) x[i] /= nrm;
    }
    // sign convention: largest component positive
    size_t imax = 0;
    for (size_t i = 1; i < m; ++i)
        if (std::fabs(x[i]) > std::fabs(x[imax])) imax = i;
    if (x[imax] < 0)
        for (size_t i = 0; i < m; ++i) x[i] = -x[i];
    *vector = x;
}

}  // namespace ofb

using namespace ofb;

// ---- operator handles --------------------------------------------------------------------------
static int linop_apply(ofb_linop *op, const void *d_in, void *d_out, cudaStream_t s) {
    switch (op->kind) {
        case 0: return launch_apply(op->plan, d_in, d_out, op->plan->n_qubits, 0, 0, op->plan->n_groups, 0, s);
        case 1: return ofb_sector_apply(op->plan, op->sector, d_in, d_out, 1, op->dim, (void *)s);
        default:
            return ofb_csr_matvec(op->dim, op->indptr, op->indices, op->data, op->index_bits, d_in, d_out,
                                  (void *)s);
    }
}

extern "C" {

int ofb_linop_from_plan(ofb_plan *plan, ofb_linop **op) {
    if (!plan || !op) return fail(OFB_ERR_INVALID, "NULL argument");
    if (plan->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    ofb_linop *o = new (std::nothrow) ofb_linop();
    if (!o) return fail(OFB_ERR_NOMEM, "host allocation failed");
    o->kind = 0;
    o->plan = plan;
    o->device = plan->device;
    o->dim = 1ll << plan->n_qubits;
    *op = o;
    return OFB_OK;
}

int ofb_linop_from_sector(ofb_plan *plan, ofb_sector *sector, ofb_linop **op) {
    if (!plan || !sector || !op) return fail(OFB_ERR_INVALID, "NULL argument");
    int64_t dim = 0;
    int rc = ofb_sector_info(sector, nullptr, &dim, nullptr);
    if (rc) return rc;
    ofb_linop *o = new (std::nothrow) ofb_linop();
    if (!o) return fail(OFB_ERR_NOMEM, "host allocation failed");
    o->kind = 1;
    o->plan = plan;
    o->sector = sector;
    o->device = plan->device;
    o->dim = dim;
    *op = o;
    return OFB_OK;
}

int ofb_linop_from_csr(int64_t n_rows, const void *d_indptr, const void *d_indices, const void *d_data,
                       int index_bits, int device, ofb_linop **op) {
    if (!op || n_rows < 0 || (index_bits != 32 && index_bits != 64)) return fail(OFB_ERR_INVALID, "bad argument");
    ofb_linop *o = new (std::nothrow) ofb_linop();
    if (!o) return fail(OFB_ERR_NOMEM, "host allocation failed");
    o->kind = 2;
    o->dim = n_rows;
    o->indptr = d_indptr;
    o->indices = d_indices;
    o->data = d_data;
    o->index_bits = index_bits;
    o->device = device;
    *op = o;
    return OFB_OK;
}

int ofb_linop_destroy(ofb_linop *op) {
    delete op;
    return OFB_OK;
}

int ofb_linop_dim(const ofb_linop *op, int64_t *dim) {
    if (!op || !dim) return fail(OFB_ERR_INVALID, "NULL argument");
    *dim = op->dim;
    return OFB_OK;
}

int ofb_linop_apply(ofb_linop *op, const void *d_in, void *d_out, void *stream) {
    if (!op || !d_in || !d_out) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(op->device));
    return linop_apply(op, d_in, d_out, (cudaStream_t)stream);
}

// ---- workspace + per-iteration primitives ---------------------------------------------------------
int ofb_kws_destroy(ofb_kws *ws) {
    if (!ws) return OFB_OK;
    cudaSetDevice(ws->device);
    cudaFree(ws->d_scal);
    cudaFree(ws->d_alpha);
    cudaFree(ws->d_beta);
    cudaFree(ws->d_part);
    if (ws->h_pin) cudaFreeHost(ws->h_pin);
    cudaGetLastError();
    delete ws;
    return OFB_OK;
}

int ofb_kws_create(int64_t n_local, int max_steps, int device, ofb_kws **out) {
    if (!out || n_local < 0 || max_steps < 1) return fail(OFB_ERR_INVALID, "bad argument");
    *out = nullptr;
    OFB_CUDA(cudaSetDevice(device));
    ofb_kws *ws = new (std::nothrow) ofb_kws();
    if (!ws) return fail(OFB_ERR_NOMEM, "host allocation failed");
    ws->device = device;
    ws->n = n_local;
    ws->max_steps = max_steps;
    cudaError_t e;
    if ((e = cudaMalloc(&ws->d_scal, SLOTS * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&ws->d_alpha, (max_steps + 1) * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&ws->d_beta, (max_steps + 2) * sizeof(double))) != cudaSuccess ||
        (e = cudaMallocHost(&ws->h_pin, (2 * (max_steps + 2) + SLOTS) * sizeof(double))) != cudaSuccess) {
        ofb_kws_destroy(ws);
        return cuda_fail(e, "workspace allocation", __FILE__, __LINE__);
    }
    int rc = kws_ensure_part(ws, 2 * (size_t)KGRID + 16);
    if (rc) {
        ofb_kws_destroy(ws);
        return rc;
    }
    *out = ws;
    return OFB_OK;
}

int ofb_kws_scalars(ofb_kws *ws, void **d_scalars) {
    if (!ws || !d_scalars) return fail(OFB_ERR_INVALID, "NULL argument");
    *d_scalars = ws->d_scal;
    return OFB_OK;
}

// start of a sweep: beta_0 (norm of the start vector, 1 for a normalised one)
int ofb_kws_reset(ofb_kws *ws, double beta0, void *stream) {
    if (!ws) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t s = (cudaStream_t)stream;
    double *h = ws->h_pin + 2 * (ws->max_steps + 2);
    OFB_CUDA(cudaStreamSynchronize(s));  // h_pin is reused
    for (int k = 0; k < SLOTS; ++k) h[k] = 0.0;
    h[S_BETA] = beta0;
    OFB_CUDA(cudaMemcpyAsync(ws->d_scal, h, SLOTS * sizeof(double), cudaMemcpyHostToDevice, s));
    OFB_CUDA(cudaMemcpyAsync(ws->d_beta, h + S_BETA, sizeof(double), cudaMemcpyHostToDevice, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    return OFB_OK;
}

// slots 0,1 = conj(x) . y over this device's n elements (no host sync)
int ofb_kws_dot(ofb_kws *ws, const void *d_x, const void *d_y, void *stream) {
    if (!ws || !d_x || !d_y) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = kgrid(ws->n);
    k_kws_dotc<<<g, KB, 0, s>>>((const double2 *)d_x, (const double2 *)d_y, ws->n, ws->d_part);
    OFB_LAUNCH_CHECK();
    return reduce_partials(ws->d_part, g, 2, ws->d_scal + S_DOT_RE, s);
}

// K1 on a group range with the dot product conj(d_dot) . out folded into the same launch
// (valid when this launch completes the output shard); result in slots 0,1
int ofb_apply_groups_dot(ofb_plan *p, const void *d_in, void *d_out, int local_bits, uint64_t index_base,
                         int64_t g_begin, int64_t g_end, int accumulate, const void *d_dot, ofb_kws *ws,
                         void *stream) {
    if (!p || !d_in || !d_out || !d_dot || !ws) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    if (local_bits < 0 || local_bits > p->n_qubits) return fail(OFB_ERR_INVALID, "bad local_bits");
    if (g_begin < 0 || g_end > p->n_groups || g_begin >= g_end) return fail(OFB_ERR_INVALID, "bad group range");
    if (d_in == d_out) return fail(OFB_ERR_INVALID, "d_in and d_out must not alias");
    OFB_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = kws_ensure_part(ws, 2 * (size_t)apply_partial_count(local_bits) + 16);
    if (rc) return rc;
    int64_t count = 0;
    rc = launch_apply_dot(p, d_in, d_out, local_bits, index_base, g_begin, g_end, accumulate, d_dot,
                          ws->d_part, &count, s);
    if (rc) return rc;
    return reduce_partials(ws->d_part, count, 2, ws->d_scal + S_DOT_RE, s);
}

int ofb_kws_alpha(ofb_kws *ws, int k, int replay, double ritz_k, void *stream) {
    if (!ws || k < 0 || k >= ws->max_steps) return fail(OFB_ERR_INVALID, "bad step");
    OFB_CUDA(cudaSetDevice(ws->device));
    k_kws_alpha<<<1, 32, 0, (cudaStream_t)stream>>>(ws->d_scal, ws->d_alpha, ws->d_beta, k, replay, ritz_k);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

// fused update: w <- w/beta - (alpha/beta) u - (beta/beta_prev) u_prev  [y += (s_k/beta) u],
// and (first pass) the local sum of |w|^2 in slot 2
int ofb_kws_step(ofb_kws *ws, void *d_w, const void *d_u, const void *d_uprev, void *d_y, int replay,
                 void *stream) {
    if (!ws || !d_w || !d_u) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = kgrid(ws->n);
    double2 *w = (double2 *)d_w, *y = (double2 *)d_y;
    const double2 *u = (const double2 *)d_u, *up = (const double2 *)d_uprev;
#define OFB_STEP(NN, RR, PP) k_lanczos_step<NN, RR, PP><<<g, KB, 0, s>>>(ws->n, w, u, up, y, ws->d_scal, ws->d_part)
    if (replay) {
        if (y) { if (up) OFB_STEP(false, true, true); else OFB_STEP(false, true, false); }
        else { if (up) OFB_STEP(false, false, true); else OFB_STEP(false, false, false); }
    } else {
        if (y) { if (up) OFB_STEP(true, true, true); else OFB_STEP(true, true, false); }
        else { if (up) OFB_STEP(true, false, true); else OFB_STEP(true, false, false); }
    }
#undef OFB_STEP
    OFB_LAUNCH_CHECK();
    if (replay) return OFB_OK;
    return reduce_partials(ws->d_part, g, 1, ws->d_scal + S_NRM2, s);
}

int ofb_kws_beta(ofb_kws *ws, int k, void *stream) {
    if (!ws || k < 0 || k >= ws->max_steps) return fail(OFB_ERR_INVALID, "bad step");
    OFB_CUDA(cudaSetDevice(ws->device));
    k_kws_beta<<<1, 32, 0, (cudaStream_t)stream>>>(ws->d_scal, ws->d_beta, k);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

// host copy of alpha_0..alpha_{count-1}, beta_0..beta_count and the breakdown flag (0 = none,
// else 1 + the step whose new vector vanished); synchronises the stream
int ofb_kws_fetch(ofb_kws *ws, int count, double *h_alphas, double *h_betas, int *h_flag, void *stream) {
    if (!ws || count < 0 || count > ws->max_steps) return fail(OFB_ERR_INVALID, "bad count");
    OFB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t s = (cudaStream_t)stream;
    double *ha = ws->h_pin, *hb = ws->h_pin + ws->max_steps + 2, *hs = ws->h_pin + 2 * (ws->max_steps + 2);
    if (count) OFB_CUDA(cudaMemcpyAsync(ha, ws->d_alpha, count * sizeof(double), cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaMemcpyAsync(hb, ws->d_beta, (count + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaMemcpyAsync(hs, ws->d_scal, SLOTS * sizeof(double), cudaMemcpyDeviceToHost, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    if (h_alphas) memcpy(h_alphas, ha, count * sizeof(double));
    if (h_betas) memcpy(h_betas, hb, (count + 1) * sizeof(double));
    if (h_flag) *h_flag = (int)hs[S_FLAG];
    return OFB_OK;
}

int ofb_tridiag_lowest(int m, const double *diag, const double *offdiag, double *value, double *vector) {
    if (m < 1 || !diag || (m > 1 && !offdiag) || !value || !vector) return fail(OFB_ERR_INVALID, "bad argument");
    OFB_TRY
    std::vector<double> d(diag, diag + m), e(offdiag, offdiag + (m - 1)), x;
    tridiag_lowest(d, e, value, &x);
    memcpy(vector, x.data(), m * sizeof(double));
    return OFB_OK;
    OFB_CATCH((void)0)
}

}  // extern "C"

// ---- the complete one-device solver ------------------------------------------------------------------
namespace {

struct DevVec {
    void *p = nullptr;
    ~DevVec() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        OFB_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return OFB_OK;
    }
};

struct Solver {
    ofb_linop *op;
    ofb_kws *ws;
    int64_t n;
    cudaStream_t s;
    int n_locked;
    const void *const *locked;
    int64_t matvecs = 0;

    int project(void *w) {  // w -= sum_l <l, w> l   (locking; host round trip, k = 2 only)
        for (int l = 0; l < n_locked; ++l) {
            double d[2];
            int rc = ofb_vec_dotc(n, locked[l], w, d, (void *)s);
            if (rc) return rc;
            double a[2] = {-d[0], -d[1]};
            rc = ofb_vec_axpy(n, a, locked[l], w, (void *)s);
            if (rc) return rc;
        }
        return OFB_OK;
    }

    // w = H u and slot 0 = <u, w>
    int apply_dot(const void *u, void *w) {
        int rc;
        ++matvecs;
        if (op->kind == 0 && n_locked == 0) {
            ofb_plan *p = op->plan;
            return ofb_apply_groups_dot(p, u, w, p->n_qubits, 0, 0, p->n_groups, 0, u, ws, (void *)s);
        }
        if ((rc = linop_apply(op, u, w, s))) return rc;
        if (n_locked && (rc = project(w))) return rc;
        return ofb_kws_dot(ws, u, w, (void *)s);
    }
};

// one sweep from `start` (unit norm).  basis != null: the u_k are kept as its columns and the
// Ritz vector is one pass over them; otherwise the recurrence is replayed with the stored
// coefficients.  Returns the number of steps m, theta, the Ritz vector in y.
int sweep(Solver &S, const void *start, int steps, double tol, void *bufA, void *bufB, void *bufC,
          void *y, void *basis, int *m_out, double *theta_out) {
    const int64_t n = S.n;
    const size_t bytes = (size_t)n * 16;
    ofb_kws *ws = S.ws;
    cudaStream_t s = S.s;
    int rc;
    auto col = [&](int k) { return (void *)((char *)basis + (size_t)k * bytes); };
    void *up = bufA, *u = bufB, *w = bufC;
    if (basis) u = col(0);
    OFB_CUDA(cudaMemcpyAsync(u, start, bytes, cudaMemcpyDeviceToDevice, s));
    if ((rc = ofb_kws_reset(ws, 1.0, (void *)s))) return rc;
    std::vector<double> alphas(steps + 1), betas(steps + 2), ritz;
    double theta = 0.0;
    int m = 0;
    for (int k = 0; k < steps; ++k) {
        if (basis) w = (k + 1 < steps) ? col(k + 1) : bufC;
        if ((rc = S.apply_dot(u, w))) return rc;
        if ((rc = ofb_kws_alpha(ws, k, 0, 0.0, (void *)s))) return rc;
        if ((rc = ofb_kws_step(ws, w, u, k ? up : nullptr, nullptr, 0, (void *)s))) return rc;
        if ((rc = ofb_kws_beta(ws, k, (void *)s))) return rc;
        m = k + 1;
        const bool check = k < 8 || (k & 3) == 3 || k == steps - 1;
        if (check) {
            int flag = 0;
            if ((rc = ofb_kws_fetch(ws, m, alphas.data(), betas.data(), &flag, (void *)s))) return rc;
            if (flag) m = std::min(m, flag);
            std::vector<double> d(alphas.begin(), alphas.begin() + m), e(m > 1 ? m - 1 : 0);
            for (int i = 0; i + 1 < m; ++i) e[i] = betas[i + 1];
            tridiag_lowest(d, e, &theta, &ritz);
            const double scale = std::max(1.0, std::fabs(theta));
            const double beta_next = betas[m];
            if (flag || std::fabs(beta_next * ritz[m - 1]) <= tol * scale || beta_next <= 1e-14 * scale) break;
        }
        if (basis) {
            up = u;
            u = w;
        } else {
            void *t = up;
            up = u;
            u = w;
            w = t;
        }
    }
    // Ritz vector y = sum_k ritz_k u_k / beta_k
    OFB_CUDA(cudaMemsetAsync(y, 0, bytes, s));
    if (basis) {
        std::vector<double> coef(2 * m);
        for (int k = 0; k < m; ++k) {
            coef[2 * k] = ritz[k] / betas[k];
            coef[2 * k + 1] = 0.0;
        }
        const double one[2] = {1.0, 0.0};
        if ((rc = ofb_vec_multi_axpy(n, m, coef.data(), basis, n, one, y, (void *)s))) return rc;
    } else {
        up = bufA; u = bufB; w = bufC;
        OFB_CUDA(cudaMemcpyAsync(u, start, bytes, cudaMemcpyDeviceToDevice, s));
        for (int k = 0; k < m; ++k) {
            if ((rc = ofb_kws_alpha(ws, k, 1, ritz[k], (void *)s))) return rc;
            if (k == m - 1) {  // the last vector only enters y
                // y += c3 * u: reuse the step kernel on a scratch w (bufA/B/C rotation leaves w free)
                if ((rc = ofb_kws_step(ws, w, u, nullptr, y, 1, (void *)s))) return rc;
                break;
            }
            ++S.matvecs;
            if ((rc = linop_apply(S.op, u, w, s))) return rc;
            if (S.n_locked && (rc = S.project(w))) return rc;
            if ((rc = ofb_kws_step(ws, w, u, k ? up : nullptr, y, 1, (void *)s))) return rc;
            void *t = up;
            up = u;
            u = w;
            w = t;
        }
    }
    *m_out = m;
    *theta_out = theta;
    return OFB_OK;
}

}  // namespace

extern "C" int ofb_lanczos_lowest(ofb_linop *op, const void *d_start, int n_locked,
                                  const void *const *d_locked, double tol, int krylov_dim,
                                  int max_restarts, uint64_t seed, int store_basis, double *h_energy,
                                  double *h_residual, void *d_vector, int64_t *h_matvecs,
                                  int *h_restarts, void *stream) {
    if (!op || !h_energy || !d_vector) return fail(OFB_ERR_INVALID, "NULL argument");
    if (n_locked < 0 || (n_locked > 0 && !d_locked)) return fail(OFB_ERR_INVALID, "bad locked list");
    OFB_TRY
    OFB_CUDA(cudaSetDevice(op->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = op->dim;
    const size_t bytes = (size_t)n * 16;
    if (krylov_dim <= 0) krylov_dim = 400;
    krylov_dim = (int)std::max<int64_t>(2, std::min<int64_t>(krylov_dim, n - n_locked));
    if (n - n_locked < 2) krylov_dim = (int)std::max<int64_t>(1, n - n_locked);
    if (max_restarts <= 0) max_restarts = 40;
    ofb_kws *ws = nullptr;
    int rc = ofb_kws_create(n, krylov_dim, op->device, &ws);
    if (rc) return rc;
    struct WsGuard { ofb_kws *w; ~WsGuard() { ofb_kws_destroy(w); } } guard{ws};
    DevVec a, b, c, y, start, hy, basis;
    if ((rc = a.alloc(bytes)) || (rc = b.alloc(bytes)) || (rc = c.alloc(bytes)) || (rc = y.alloc(bytes)) ||
        (rc = start.alloc(bytes)) || (rc = hy.alloc(bytes)))
        return rc;
    int stored = 0;
    if (store_basis) {
        size_t free_b = 0, total_b = 0;
        OFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const int64_t fit = (int64_t)(0.7 * (double)free_b) / (int64_t)std::max<size_t>(bytes, 16);
        if (fit >= 16 || fit >= krylov_dim) stored = (int)std::min<int64_t>(krylov_dim, fit);
        if (stored && basis.alloc((size_t)stored * bytes)) {
            cudaGetLastError();
            stored = 0;
        }
    }
    if (d_start) {
        OFB_CUDA(cudaMemcpyAsync(start.p, d_start, bytes, cudaMemcpyDeviceToDevice, s));
    } else if ((rc = ofb_vec_fill_random(n, seed, 0, start.p, (void *)s))) {
        return rc;
    }
    Solver S{op, ws, n, s, n_locked, d_locked};
    double theta = 0.0, residual = INFINITY;
    int restart = 0;
    for (; restart < max_restarts; ++restart) {
        if (n_locked && (rc = S.project(start.p))) return rc;
        double norm = 0.0;
        if ((rc = ofb_vec_nrm2(n, start.p, &norm, (void *)s))) return rc;
        if (!(norm > 0.0)) {
            if ((rc = ofb_vec_fill_random(n, seed + 7919ull * (uint64_t)(restart + 1), 0, start.p, (void *)s))) return rc;
            continue;
        }
        const double inv[2] = {1.0 / norm, 0.0};
        if ((rc = ofb_vec_scal(n, inv, start.p, (void *)s))) return rc;
        int m = 0;
        const int steps = stored ? stored : krylov_dim;
        if ((rc = sweep(S, start.p, steps, tol, a.p, b.p, c.p, y.p, stored ? basis.p : nullptr, &m, &theta)))
            return rc;
        if (n_locked && (rc = S.project(y.p))) return rc;
        double ynorm = 0.0;
        if ((rc = ofb_vec_nrm2(n, y.p, &ynorm, (void *)s))) return rc;
        const double yinv[2] = {1.0 / ynorm, 0.0};
        if ((rc = ofb_vec_scal(n, yinv, y.p, (void *)s))) return rc;
        // true residual |H y - theta y| with theta the Rayleigh quotient
        ++S.matvecs;
        if ((rc = linop_apply(op, y.p, hy.p, s))) return rc;
        double d[2];
        if ((rc = ofb_vec_dotc(n, y.p, hy.p, d, (void *)s))) return rc;
        theta = d[0];
        const double mt[2] = {-theta, 0.0};
        if ((rc = ofb_vec_axpy(n, mt, y.p, hy.p, (void *)s))) return rc;
        if ((rc = ofb_vec_nrm2(n, hy.p, &residual, (void *)s))) return rc;
        OFB_CUDA(cudaMemcpyAsync(start.p, y.p, bytes, cudaMemcpyDeviceToDevice, s));
        if (residual <= 10.0 * tol * std::max(1.0, std::fabs(theta))) {
            ++restart;
            break;
        }
    }
    OFB_CUDA(cudaMemcpyAsync(d_vector, start.p, bytes, cudaMemcpyDeviceToDevice, s));
    OFB_CUDA(cudaStreamSynchronize(s));
    *h_energy = theta;
    if (h_residual) *h_residual = residual;
    if (h_matvecs) *h_matvecs = S.matvecs;
    if (h_restarts) *h_restarts = restart;
    return OFB_OK;
    OFB_CATCH((void)0)
}
