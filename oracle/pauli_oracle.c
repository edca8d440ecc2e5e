/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (plain C restatement of the reference matvec).
 *
 * Restates LinearQubitOperator._matvec
 * (/root/reference/src/openfermion/linalg/linear_qubit_operator.py:107-148) term by
 * term in the reference's scatter form and accumulation order:
 *     for each term t (dict order):
 *         amplitudes = (coeff_t * i^{y_t}) * vec            (:136)
 *         amplitudes = signs * amplitudes                    (:141-142), signs = 1 - 2*parity(i & z)
 *         out[i ^ x_t] += amplitudes[i]                      (:143-146)
 * Each output element therefore receives its contributions in term order, exactly as
 * numpy does; the loop over i inside one term is parallel (distinct targets).  Used by
 * tests/ and bench.py's cpu_baseline leg for sizes where the numpy port takes minutes.
 * Parity pinning: tests/test_oracle_golden.py checks it against the numpy port
 * (oracle/pauli_oracle.py), which is itself checked against the reference's known-answer
 * tests and against outputs of the real reference (tests/golden/).
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (see __graft_entry__.build).
 */
#include <stdint.h>
#include <string.h>

static inline unsigned parity64(uint64_t v) { return (unsigned)__builtin_parityll(v); }

/* x, z: index-convention masks; coeff_ri: (re, im) of the reference coefficient;
 * vec/out: interleaved complex128 of length 2^n. Returns 0. */
int oracle_matvec(int n_qubits, int64_t n_terms, const uint64_t *x, const uint64_t *z,
                  const double *coeff_ri, const double *vec, double *out) {
    const int64_t dim = (int64_t)1 << n_qubits;
    memset(out, 0, (size_t)dim * 2 * sizeof(double));
    for (int64_t t = 0; t < n_terms; ++t) {
        const uint64_t xm = x[t], zm = z[t];
        const int y = __builtin_popcountll(xm & zm) & 3;
        /* coefficient * (1, 1j, -1, -1j)[y % 4] as Python complex multiplication */
        static const double pr[4] = {1.0, 0.0, -1.0, -0.0}, pi[4] = {0.0, 1.0, 0.0, -1.0};
        const double ar = coeff_ri[2 * t], ai = coeff_ri[2 * t + 1];
        const double cr = ar * pr[y] - ai * pi[y];
        const double ci = ar * pi[y] + ai * pr[y];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dim; ++i) {
            const double vr = vec[2 * i], vi = vec[2 * i + 1];
            double tr = cr * vr - ci * vi;
            double ti = cr * vi + ci * vr;
            if (zm && parity64((uint64_t)i & zm)) {
                tr = -tr;
                ti = -ti;
            }
            const int64_t j = i ^ (int64_t)xm;
            out[2 * j] += tr;
            out[2 * j + 1] += ti;
        }
    }
    return 0;
}

/* Diagonal restatement (sparse_tools.py:197-247): only strings without X/Y contribute. */
int oracle_diagonal(int n_qubits, int64_t n_terms, const uint64_t *x, const uint64_t *z,
                    const double *coeff_ri, double *out) {
    const int64_t dim = (int64_t)1 << n_qubits;
    memset(out, 0, (size_t)dim * sizeof(double));
    for (int64_t t = 0; t < n_terms; ++t) {
        if (x[t]) continue;
        const uint64_t zm = z[t];
        const double c = coeff_ri[2 * t];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dim; ++i) out[i] += parity64((uint64_t)i & zm) ? -c : c;
    }
    return 0;
}
