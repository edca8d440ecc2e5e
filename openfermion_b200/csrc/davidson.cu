// Davidson's method behind the C ABI (linalg/davidson.py:99-339 of the reference: get_lowest_n,
// _iterate, _get_new_directions, orthonormalize, append_random_vectors).
//
// The reference runs the loop in numpy on 2^n-long host arrays; the Python mirror
// (openfermion_b200/linalg/davidson.py) keeps the vectors on the device but still drives every
// kernel through ctypes and hands the projected matrix to numpy.linalg.eigh.  Here the whole loop
// is native: guess_v / guess_mv are two device column blocks, the operator is an ofb_linop (K1,
// K7 or CSR), the m x m projected problem is solved on the host by a Householder reduction of the
// Hermitian matrix to real tridiagonal form followed by implicit QL (what LAPACK's zheev does),
// and only its scalars cross the bus.  Same iteration structure, same drop rules (largest |entry|
// below eps), same restart (Ritz vectors + new directions), same clamped-inverse-diagonal
// correction (ofb_davidson_correction) as the reference; random vectors come from the library's
// counter-based generator (ofb_vec_fill_random) instead of numpy.random.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

#include "ofb_internal.cuh"

using namespace ofb;
typedef std::complex<double> cplx;

namespace {

// ---- Hermitian eigenproblem of the projected matrix ---------------------------------------------
// a: column-major m x m, only the lower triangle (i >= j) is read.  w ascending, z column-major
// orthonormal eigenvectors.  Householder tridiagonalisation (zhetd2-style rank-2 updates), Q
// accumulated explicitly, then implicit QL with Wilkinson shifts on the real tridiagonal matrix
// with the rotations applied to the complex Q.
int hermitian_eigh(int m, const cplx *a_in, double *w, cplx *z) {
    if (m <= 0) return 0;
    std::vector<cplx> A((size_t)m * m);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i)
            A[(size_t)j * m + i] = i >= j ? a_in[(size_t)j * m + i] : std::conj(a_in[(size_t)i * m + j]);
    auto at = [&](int i, int j) -> cplx & { return A[(size_t)j * m + i]; };
    std::vector<double> d(m), e(m, 0.0);
    std::vector<cplx> tau(std::max(m - 1, 1)), x(m), v(m);
    for (int i = 0; i + 1 < m; ++i) {
        // reflector H = I - tau v v^H (v[0] = 1) with H^H A[i+1:, i] = beta e_1, beta real
        const int len = m - i - 1;
        const cplx alpha = at(i + 1, i);
        double xnorm2 = 0.0;
        for (int k = 1; k < len; ++k) xnorm2 += std::norm(at(i + 1 + k, i));
        if (xnorm2 == 0.0 && alpha.imag() == 0.0) {
            tau[i] = 0.0;
            e[i] = alpha.real();
            continue;
        }
        double beta = std::sqrt(std::norm(alpha) + xnorm2);
        if (alpha.real() > 0.0) beta = -beta;
        tau[i] = cplx((beta - alpha.real()) / beta, -alpha.imag() / beta);
        const cplx scale = 1.0 / (alpha - beta);
        v[0] = 1.0;
        for (int k = 1; k < len; ++k) v[k] = at(i + 1 + k, i) * scale;
        e[i] = beta;
        // x = tau A22 v
        for (int r = 0; r < len; ++r) {
            cplx acc = 0.0;
            for (int c = 0; c < len; ++c) {
                const cplx arc = r >= c ? at(i + 1 + r, i + 1 + c) : std::conj(at(i + 1 + c, i + 1 + r));
                acc += arc * v[c];
            }
            x[r] = tau[i] * acc;
        }
        cplx xv = 0.0;  // x^H v
        for (int r = 0; r < len; ++r) xv += std::conj(x[r]) * v[r];
        const cplx al = -0.5 * tau[i] * xv;
        for (int r = 0; r < len; ++r) x[r] += al * v[r];  // w
        for (int c = 0; c < len; ++c)
            for (int r = c; r < len; ++r)
                at(i + 1 + r, i + 1 + c) -= v[r] * std::conj(x[c]) + x[r] * std::conj(v[c]);
        // keep v below the subdiagonal for the accumulation of Q
        at(i + 1, i) = 1.0;
        for (int k = 1; k < len; ++k) at(i + 1 + k, i) = v[k];
    }
    for (int i = 0; i < m; ++i) d[i] = at(i, i).real();
    // Q = H_0 H_1 ... H_{m-2}, built backwards into z
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) z[(size_t)j * m + i] = i == j ? 1.0 : 0.0;
    for (int i = m - 2; i >= 0; --i) {
        if (tau[i] == cplx(0.0)) continue;
        const int len = m - i - 1;
        // z[i+1:, i+1:] <- (I - tau v v^H) z[i+1:, i+1:]
        for (int c = i + 1; c < m; ++c) {
            cplx dot = 0.0;
            for (int k = 0; k < len; ++k) dot += std::conj(at(i + 1 + k, i)) * z[(size_t)c * m + i + 1 + k];
            dot *= tau[i];
            for (int k = 0; k < len; ++k) z[(size_t)c * m + i + 1 + k] -= at(i + 1 + k, i) * dot;
        }
    }
    // implicit QL (EISPACK tql2 / Numerical Recipes tqli) on (d, e), rotations applied to z
    for (int l = 0; l < m; ++l) {
        int iter = 0, mm;
        do {
            for (mm = l; mm + 1 < m; ++mm) {
                const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
                if (std::fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm != l) {
                if (++iter > 60) return 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = mm - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    r = std::hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    cplx *zi = z + (size_t)i * m, *zi1 = z + (size_t)(i + 1) * m;
                    for (int k = 0; k < m; ++k) {
                        const cplx fz = zi1[k];
                        zi1[k] = s * zi[k] + c * fz;
                        zi[k] = c * zi[k] - s * fz;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[mm] = 0.0;
            }
        } while (mm != l);
    }
    // ascending order (selection sort with column swaps)
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (d[j] < d[k]) k = j;
        if (k != i) {
            std::swap(d[i], d[k]);
            for (int r = 0; r < m; ++r) std::swap(z[(size_t)i * m + r], z[(size_t)k * m + r]);
        }
        w[i] = d[i];
    }
    return 0;
}

// ---- device column blocks -----------------------------------------------------------------------
struct Block {
    char *buf = nullptr;
    int64_t dim = 0;
    int capacity = 0, count = 0;
    void *col(int k) const { return buf + (size_t)k * dim * 16; }
};

struct Ctx {
    ofb_linop *op;
    int64_t dim;
    cudaStream_t s;
    uint64_t seed;
    int real_only;
};

#define DV(call)             \
    do {                     \
        int rc__ = (call);   \
        if (rc__) return rc__; \
    } while (0)

int nrm2(const Ctx &c, const void *x, double *out) { return ofb_vec_nrm2(c.dim, x, out, (void *)c.s); }
int maxabs(const Ctx &c, const void *x, double *out) { return ofb_vec_maxabs(c.dim, x, out, (void *)c.s); }
int copy(const Ctx &c, const void *src, void *dst) {
    OFB_CUDA(cudaMemcpyAsync(dst, src, (size_t)c.dim * 16, cudaMemcpyDeviceToDevice, c.s));
    return OFB_OK;
}

// largest |Im x_j| (through the scratch column R): the reference's numpy.allclose(real(v), v) test
int max_imag(const Ctx &c, const void *x, void *scratch, double *out) {
    DV(copy(c, x, scratch));
    DV(ofb_vec_drop_imag(c.dim, scratch, (void *)c.s));
    const double minus[2] = {-1.0, 0.0};
    DV(ofb_vec_axpy(c.dim, minus, x, scratch, (void *)c.s));  // Re x - x = -i Im x
    return maxabs(c, scratch, out);
}

// Gram-Schmidt of columns [first_new, count) against the kept columns before them, twice; a column
// whose largest |entry| falls below eps (times its norm when `relative`) is dropped
// (linalg/davidson.py:439-469), survivors are packed.
int orthonormalize(const Ctx &c, Block &b, int first_new, double eps, bool relative) {
    int kept = first_new;
    std::vector<double> proj;
    const double one[2] = {1.0, 0.0};
    for (int i = first_new; i < b.count; ++i) {
        void *vi = b.col(i);
        double scale = 1.0;
        if (relative) DV(nrm2(c, vi, &scale));
        for (int round = 0; round < 2 && kept; ++round) {
            proj.resize((size_t)2 * kept);
            DV(ofb_vec_multi_dotc(c.dim, kept, b.col(0), c.dim, vi, proj.data(), (void *)c.s));
            for (double &t : proj) t = -t;
            DV(ofb_vec_multi_axpy(c.dim, kept, proj.data(), b.col(0), c.dim, one, vi, (void *)c.s));
        }
        double big = 0.0;
        DV(maxabs(c, vi, &big));
        if (big < eps * scale) continue;
        double norm = 0.0;
        DV(nrm2(c, vi, &norm));
        const double inv[2] = {1.0 / norm, 0.0};
        DV(ofb_vec_scal(c.dim, inv, vi, (void *)c.s));
        if (kept != i) DV(copy(c, vi, b.col(kept)));
        ++kept;
    }
    b.count = kept;
    return OFB_OK;
}

// append_random_vectors (linalg/davidson.py:390-436) on a device block; *warned = 1 when fewer
// columns than asked for could be produced
int append_random(Ctx &c, Block &b, int cols, int *warned) {
    if (cols <= 0) return OFB_OK;
    int have = b.count;
    const int64_t want64 = std::min<int64_t>(std::min<int64_t>((int64_t)have + cols, c.dim + 1), b.capacity);
    const int want = (int)want64;
    int trial = 0;
    while (have < want) {
        ++trial;
        for (int k = have; k < want; ++k)
            DV(ofb_vec_fill_random(c.dim, c.seed++, c.real_only, b.col(k), (void *)c.s));
        b.count = want;
        DV(orthonormalize(c, b, have, 1e-6, false));
        if (b.count == have) {
            if (trial > 3) {
                if (warned) *warned = 1;
                break;
            }
        } else {
            trial = 1;
            have = b.count;
        }
    }
    return OFB_OK;
}

}  // namespace

extern "C" {

int ofb_hermitian_eigh(int m, const double *a_ri, double *w, double *z_ri) {
    if (m < 0 || (m > 0 && (!a_ri || !w || !z_ri))) return fail(OFB_ERR_INVALID, "bad argument");
    OFB_TRY
    if (hermitian_eigh(m, (const cplx *)a_ri, w, (cplx *)z_ri))
        return fail(OFB_ERR_STATE, "QL iteration did not converge");
    return OFB_OK;
    OFB_CATCH((void)0)
}

int ofb_davidson_lowest_n(ofb_linop *op, const double *d_diag, int n_lowest, int n_guess,
                          const void *d_guess, int max_subspace, int max_iterations, double eps,
                          int real_only, uint64_t seed, double *h_eigenvalues, void *d_eigenvectors,
                          int *h_success, int *h_iterations, double *h_max_error, int *h_flags,
                          void *stream) {
    if (!op || !d_diag || !h_eigenvalues || !d_eigenvectors)
        return fail(OFB_ERR_INVALID, "NULL argument");
    int64_t dim = 0;
    DV(ofb_linop_dim(op, &dim));
    if (dim <= 0) return fail(OFB_ERR_INVALID, "empty operator");
    if (max_subspace <= 2 || max_iterations <= 0 || !(eps > 0.0))
        return fail(OFB_ERR_INVALID, "Invalid values for max_subspace, max_iterations and/ or eps");
    max_subspace = (int)std::min<int64_t>(max_subspace, dim + 1);  // DavidsonOptions.set_dimension
    if (n_lowest <= 0 || n_lowest >= max_subspace)
        return fail(OFB_ERR_INVALID, "n_lowest is supposed to be in [1, max_subspace)");
    if (n_guess < 0 || (n_guess > 0 && !d_guess)) return fail(OFB_ERR_INVALID, "bad initial guess");
    int device = 0;
    OFB_CUDA(cudaGetDevice(&device));
    Ctx c{op, dim, (cudaStream_t)stream, seed, real_only};
    if (h_flags) *h_flags = 0;

    // work space: guess_v / guess_mv blocks that fit in 45 % of the free memory
    size_t free_b = 0, total_b = 0;
    OFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const int64_t fit = (int64_t)(0.45 * (double)free_b / ((double)dim * 16.0)) - 2 * n_lowest - 2;
    int capacity = (int)std::max<int64_t>(std::min<int64_t>((int64_t)max_subspace + n_lowest, fit),
                                          n_lowest + 2);
    capacity = std::max(capacity, std::max(n_guess, n_lowest) + n_lowest);
    max_subspace = std::min(max_subspace, capacity - n_lowest);

    Block V, MV, TV, TMV;
    void *R = nullptr;
    auto release = [&]() {
        cudaFree(V.buf);
        cudaFree(MV.buf);
        cudaFree(TV.buf);
        cudaFree(TMV.buf);
        cudaFree(R);
    };
    auto alloc = [&](Block &b, int cap) -> int {
        b.dim = dim;
        b.capacity = cap;
        OFB_CUDA(cudaMalloc((void **)&b.buf, (size_t)cap * dim * 16));
        return OFB_OK;
    };
    int rc = alloc(V, capacity);
    if (!rc) rc = alloc(MV, capacity);
    if (!rc) rc = alloc(TV, n_lowest);
    if (!rc) rc = alloc(TMV, n_lowest);
    if (!rc && cudaMalloc(&R, (size_t)dim * 16) != cudaSuccess) rc = fail(OFB_ERR_NOMEM, "device allocation failed");
    if (rc) {
        release();
        return rc;
    }

    std::vector<cplx> vmv, small, zvec;
    std::vector<double> lam, row, coef;
    int vmv_valid = 0;
    int success = 0, iterations = 0;
    double max_trial_error = 0.0;
    int k_out = 0;
    const double zero[2] = {0.0, 0.0};

    auto body = [&]() -> int {
        vmv.assign((size_t)capacity * capacity, cplx(0.0));
        if (n_guess > 0) {
            OFB_CUDA(cudaMemcpyAsync(V.buf, d_guess, (size_t)n_guess * dim * 16, cudaMemcpyDeviceToDevice, c.s));
            V.count = n_guess;
            if (real_only)
                for (int k = 0; k < n_guess; ++k) DV(ofb_vec_drop_imag(dim, V.col(k), (void *)c.s));
            // scipy.linalg.orth of the reference (davidson.py:154): an orthonormal basis of the span
            DV(orthonormalize(c, V, 0, 1e-12, true));
        }
        if (V.count < n_lowest) DV(append_random(c, V, n_lowest - V.count, h_flags));
        if (V.count == 0) return fail(OFB_ERR_INVALID, "Guess vectors are all zero!");

        while (iterations < max_iterations && !success) {
            // ---- one iteration (davidson.py:223-286) ----
            for (int k = MV.count; k < V.count; ++k)
                DV(ofb_linop_apply(op, V.col(k), MV.col(k), (void *)c.s));
            MV.count = V.count;
            const int m = V.count;
            // projected matrix, lower triangle: vmv[i][j] = v_i^H mv_j (i >= j), new rows only
            for (int i = vmv_valid; i < m; ++i) {
                row.resize((size_t)2 * (i + 1));
                DV(ofb_vec_multi_dotc(dim, i + 1, MV.col(0), dim, V.col(i), row.data(), (void *)c.s));
                for (int j = 0; j <= i; ++j)  // row[j] = mv_j^H v_i
                    vmv[(size_t)j * capacity + i] = cplx(row[2 * j], -row[2 * j + 1]);
            }
            vmv_valid = m;
            small.resize((size_t)m * m);
            zvec.resize((size_t)m * m);
            lam.resize(m);
            for (int j = 0; j < m; ++j)
                for (int i = 0; i < m; ++i) small[(size_t)j * m + i] = vmv[(size_t)j * capacity + i];
            if (hermitian_eigh(m, small.data(), lam.data(), zvec.data()))
                return fail(OFB_ERR_STATE, "projected eigenproblem did not converge");
            k_out = std::min(n_lowest, m);
            // Ritz vectors and their images
            for (int i = 0; i < k_out; ++i) {
                const double *t = (const double *)(zvec.data() + (size_t)i * m);
                DV(ofb_vec_multi_axpy(dim, m, t, V.col(0), dim, zero, TV.col(i), (void *)c.s));
                DV(ofb_vec_multi_axpy(dim, m, t, MV.col(0), dim, zero, TMV.col(i), (void *)c.s));
            }
            TV.count = TMV.count = k_out;
            for (int i = 0; i < k_out; ++i) h_eigenvalues[i] = lam[i];
            // residuals and new directions (davidson.py:288-339)
            max_trial_error = 0.0;
            for (int i = 0; i < k_out; ++i) {
                DV(copy(c, TMV.col(i), R));
                const double ml[2] = {-lam[i], 0.0};
                DV(ofb_vec_axpy(dim, ml, TV.col(i), R, (void *)c.s));
                double big = 0.0;
                DV(maxabs(c, R, &big));
                if (big < eps) continue;
                double nr = 0.0;
                DV(nrm2(c, R, &nr));
                max_trial_error = std::max(max_trial_error, nr);
                if (V.count >= V.capacity) return fail(OFB_ERR_STATE, "Davidson work space exhausted; lower max_subspace.");
                DV(ofb_davidson_correction(dim, d_diag, lam[i], eps, R, TV.col(i), V.col(V.count), (void *)c.s));
                V.count += 1;
            }
            if (max_trial_error < eps) {
                success = 1;
                break;
            }
            if (real_only) {
                for (int k = 0; k < V.count; ++k) DV(ofb_vec_drop_imag(dim, V.col(k), (void *)c.s));
                vmv_valid = 0;
            }
            const int count_mvs = MV.count;
            DV(orthonormalize(c, V, count_mvs, eps, false));
            if (V.count <= count_mvs) DV(append_random(c, V, n_lowest, h_flags));
            if (V.count >= max_subspace) {
                // restart: Ritz vectors + new directions (davidson.py:201-211)
                const int new_count = V.count - count_mvs;
                for (int k = 0; k < new_count; ++k) {
                    DV(copy(c, V.col(count_mvs + k), R));
                    DV(copy(c, R, V.col(TV.count + k)));
                }
                for (int k = 0; k < TV.count; ++k) {
                    DV(copy(c, TV.col(k), V.col(k)));
                    DV(copy(c, TMV.col(k), MV.col(k)));
                }
                V.count = TV.count + new_count;
                MV.count = TV.count;
                vmv_valid = 0;
                if (real_only) {  // complex Ritz vectors: their images are recomputed (davidson.py:204-210)
                    double worst = 0.0, im = 0.0;
                    for (int k = 0; k < V.count; ++k) {
                        DV(max_imag(c, V.col(k), R, &im));
                        worst = std::max(worst, im);
                    }
                    for (int k = 0; k < MV.count; ++k) {
                        DV(max_imag(c, MV.col(k), R, &im));
                        worst = std::max(worst, im);
                    }
                    if (worst > 1e-8) MV.count = 0;
                }
            }
            ++iterations;
        }
        OFB_CUDA(cudaMemcpyAsync(d_eigenvectors, TV.buf, (size_t)k_out * dim * 16, cudaMemcpyDeviceToDevice, c.s));
        OFB_CUDA(cudaStreamSynchronize(c.s));
        return OFB_OK;
    };
    OFB_TRY
    rc = body();
    OFB_CATCH(release())
    release();
    if (h_success) *h_success = success;
    if (h_iterations) *h_iterations = iterations;
    if (h_max_error) *h_max_error = max_trial_error;
    return rc;
}

}  // extern "C"
