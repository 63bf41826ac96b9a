// TEST INFRASTRUCTURE ONLY — the extended-precision parity target (liboracle_exact.so).  Nothing under plssvm_b200/ may include, link or call this.
//
// The reference's CPU path is not reproducible run to run (atomics in src/plssvm/backends/OpenMP/svm_kernel.cpp:45-51) and CG started from
// x0 = 1 amplifies rounding noise, so "repo vs reference" can only be compared up to the reference's own spread.  This file restates the SAME
// algorithm — kernel functions include/plssvm/kernel_function_types.hpp:75-97, q vector OpenMP/q_kernel.cpp:18-52, implicit matvec
// OpenMP/svm_kernel.cpp:22-55, CG driver OpenMP/csvm.cpp:71-183 (identical to gpu_csvm.hpp:477-654), predict OpenMP/csvm.cpp:188-227 — in
// arithmetic far beyond either implementation, deterministically (every output element is produced by one thread in a fixed order):
//   * inner products / squared distances: compensated (Ogita-Rump-Oishi Dot2, error-free TwoProd / TwoSum) — as accurate as twice the working
//     precision (~106 bits), evaluated on the inputs exactly as given (float or double);
//   * kernel function, matvec sums, CG vector algebra and scalars: x87 long double (64-bit mantissa).
// With it both sides get an error |. - exact| instead of a difference to a noisy partner (tests/parity.py, profiles/r02/parity_report.json).
// Pinned by tests/test_oracle_exact.py: it reproduces the reference's known-answer fixtures and agrees with the reference build to the
// reference's own rounding level.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {

using ld = long double;

inline void two_sum(const double a, const double b, double &s, double &e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
inline void two_prod(const double a, const double b, double &p, double &e) {
    p = a * b;
    e = std::fma(a, b, -p);
}

// sum_k x_k y_k with twice the working precision
template <typename T>
ld dot2(const T *x, const T *y, const std::size_t d) {
    double s = 0.0, c = 0.0;
    for (std::size_t k = 0; k < d; ++k) {
        double p, e1, e2;
        two_prod(static_cast<double>(x[k]), static_cast<double>(y[k]), p, e1);
        two_sum(s, p, s, e2);
        c += e1 + e2;
    }
    return static_cast<ld>(s) + static_cast<ld>(c);
}
// sum_k (x_k - y_k)^2 with twice the working precision (the difference itself is kept exactly as hi + lo)
template <typename T>
ld dist2(const T *x, const T *y, const std::size_t d) {
    double s = 0.0, c = 0.0;
    for (std::size_t k = 0; k < d; ++k) {
        double dh, dl, p, e1, e2;
        two_sum(static_cast<double>(x[k]), -static_cast<double>(y[k]), dh, dl);
        two_prod(dh, dh, p, e1);
        e1 += 2.0 * dh * dl;
        two_sum(s, p, s, e2);
        c += e1 + e2;
    }
    return static_cast<ld>(s) + static_cast<ld>(c);
}

template <typename T>
struct params {
    int kernel, degree;
    T gamma, coef0;
};

// kernel_function_types.hpp:75-97
template <typename T>
ld kernel_function(const T *x, const T *y, const std::size_t d, const params<T> &p) {
    switch (p.kernel) {
        case 0: return dot2(x, y, d);
        case 1: return std::pow(static_cast<ld>(p.gamma) * dot2(x, y, d) + static_cast<ld>(p.coef0), static_cast<ld>(p.degree));
        default: return std::exp(-static_cast<ld>(p.gamma) * dist2(x, y, d));
    }
}

// rows `rows[0..n_rows)` of ret = Q~ v, Q~_ij = k(x_i, x_j) + QA_cost - q_i - q_j + delta_ij cost_inv (svm_kernel.cpp:22-55 without the mirror trick)
template <typename T>
void matvec_rows(const T *X, const std::size_t N, const std::size_t d, const ld *q, const ld *v, const ld QA_cost, const ld cost_inv, const params<T> &p, const std::uint64_t *rows,
                 const std::size_t n_rows, ld *out) {
    const std::size_t n = N - 1;
    #pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t r = 0; r < n_rows; ++r) {
        const std::size_t i = rows == nullptr ? r : static_cast<std::size_t>(rows[r]);
        ld s = 0.0L;
        for (std::size_t j = 0; j < n; ++j) {
            ld t = kernel_function(X + i * d, X + j * d, d, p) + QA_cost - q[i] - q[j];
            if (i == j) { t += cost_inv; }
            s += t * v[j];
        }
        out[r] = s;
    }
}

template <typename T>
struct solver {
    const T *X;
    std::size_t N, d, n;
    params<T> p;
    std::vector<ld> K;  // n x n kernel matrix incl. the rank structure, computed once
    std::vector<ld> q;
    ld QA_cost, cost_inv;

    solver(const T *X_, const std::size_t N_, const std::size_t d_, const params<T> &p_, const T cost) : X(X_), N(N_), d(d_), n(N_ - 1), p(p_), K(n * n), q(n) {
        cost_inv = 1.0L / static_cast<ld>(cost);
        for (std::size_t i = 0; i < n; ++i) { q[i] = kernel_function(X + i * d, X + n * d, d, p); }  // q_kernel.cpp:18-52
        QA_cost = kernel_function(X + n * d, X + n * d, d, p) + cost_inv;                           // csvm.cpp:86
        #pragma omp parallel for schedule(dynamic, 4)
        for (std::size_t i = 0; i < n; ++i) {
            for (std::size_t j = 0; j <= i; ++j) {
                ld t = kernel_function(X + i * d, X + j * d, d, p) + QA_cost - q[i] - q[j];
                if (i == j) { t += cost_inv; }
                K[i * n + j] = t;
            }
        }
        for (std::size_t i = 0; i < n; ++i) {
            for (std::size_t j = i + 1; j < n; ++j) { K[i * n + j] = K[j * n + i]; }
        }
    }
    void apply(const std::vector<ld> &v, std::vector<ld> &out) const {
        #pragma omp parallel for
        for (std::size_t i = 0; i < n; ++i) {
            ld s = 0.0L;
            for (std::size_t j = 0; j < n; ++j) { s += K[i * n + j] * v[j]; }
            out[i] = s;
        }
    }
};

ld vdot(const std::vector<ld> &a, const std::vector<ld> &b) {
    ld s = 0.0L;
    for (std::size_t i = 0; i < a.size(); ++i) { s += a[i] * b[i]; }
    return s;
}

// csvm.cpp:71-183, every quantity in long double
template <typename T>
int exact_solve(const int kernel, const T *X, const std::size_t N, const std::size_t d, const T *y, const int degree, const T gamma, const T coef0, const T cost, const T eps,
                const std::uint64_t max_iter, double *alpha_out, double *rho_out, std::uint64_t *iters_out, double *trace /* optional, max_iter + 1 */) {
    if (N < 2 || d == 0 || !(eps > T{ 0 }) || max_iter == 0) { return 1; }
    const solver<T> A(X, N, d, params<T>{ kernel, degree, gamma, coef0 }, cost);
    const std::size_t n = N - 1;
    std::vector<ld> b(n), x(n, 1.0L), r(n), dvec(n), Ad(n);
    for (std::size_t i = 0; i < n; ++i) { b[i] = static_cast<ld>(y[i]) - static_cast<ld>(y[n]); }
    A.apply(x, Ad);
    for (std::size_t i = 0; i < n; ++i) { r[i] = b[i] - Ad[i]; }
    ld delta = vdot(r, r);
    const ld delta0 = delta;
    dvec = r;
    if (trace != nullptr) { trace[0] = static_cast<double>(delta0); }
    const ld eps2 = static_cast<ld>(eps) * static_cast<ld>(eps);
    std::uint64_t iter = 0;
    for (; iter < max_iter; ++iter) {
        A.apply(dvec, Ad);
        const ld alpha_cd = delta / vdot(dvec, Ad);
        for (std::size_t i = 0; i < n; ++i) { x[i] += alpha_cd * dvec[i]; }
        if (iter % 50 == 49) {
            A.apply(x, Ad);
            for (std::size_t i = 0; i < n; ++i) { r[i] = b[i] - Ad[i]; }
        } else {
            for (std::size_t i = 0; i < n; ++i) { r[i] -= alpha_cd * Ad[i]; }
        }
        const ld delta_old = delta;
        delta = vdot(r, r);
        if (trace != nullptr) { trace[iter + 1] = static_cast<double>(delta); }
        if (delta <= eps2 * delta0) { break; }
        const ld beta = delta / delta_old;
        for (std::size_t i = 0; i < n; ++i) { dvec[i] = beta * dvec[i] + r[i]; }
    }
    ld sum = 0.0L, qx = 0.0L;
    for (std::size_t i = 0; i < n; ++i) {
        sum += x[i];
        qx += A.q[i] * x[i];
        alpha_out[i] = static_cast<double>(x[i]);
    }
    alpha_out[n] = static_cast<double>(-sum);
    *rho_out = static_cast<double>(-(static_cast<ld>(y[n]) + A.QA_cost * sum - qx));
    if (iters_out != nullptr) { *iters_out = std::min<std::uint64_t>(iter + 1, max_iter); }
    return 0;
}

template <typename T>
void exact_matvec(const int kernel, const T *X, const std::size_t N, const std::size_t d, const T *q, const T *v, const T QA_cost, const T cost_inv, const int degree, const T gamma,
                  const T coef0, const std::uint64_t *rows, const std::size_t n_rows, double *out) {
    const std::size_t n = N - 1;
    std::vector<ld> ql(n), vl(n), res(rows == nullptr ? n : n_rows);
    for (std::size_t i = 0; i < n; ++i) {
        ql[i] = static_cast<ld>(q[i]);
        vl[i] = static_cast<ld>(v[i]);
    }
    matvec_rows(X, N, d, ql.data(), vl.data(), static_cast<ld>(QA_cost), static_cast<ld>(cost_inv), params<T>{ kernel, degree, gamma, coef0 }, rows, res.size(), res.data());
    for (std::size_t r = 0; r < res.size(); ++r) { out[r] = static_cast<double>(res[r]); }
}

template <typename T>
void exact_q(const int kernel, const T *X, const std::size_t N, const std::size_t d, const int degree, const T gamma, const T coef0, double *q_out /* N: entry N-1 = k(x_N, x_N) */) {
    const params<T> p{ kernel, degree, gamma, coef0 };
    #pragma omp parallel for
    for (std::size_t i = 0; i < N; ++i) { q_out[i] = static_cast<double>(kernel_function(X + i * d, X + (N - 1) * d, d, p)); }
}

// csvm.cpp:188-227 (the linear kernel is evaluated as sum_i alpha_i <sv_i, p> as well — the same value as <w, p> in exact arithmetic)
template <typename T>
void exact_predict(const int kernel, const T *SV, const std::size_t n_sv, const std::size_t d, const T *alpha, const T rho, const T *P, const std::size_t m, const int degree,
                   const T gamma, const T coef0, double *out) {
    const params<T> p{ kernel, degree, gamma, coef0 };
    #pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t pt = 0; pt < m; ++pt) {
        ld s = 0.0L;
        for (std::size_t i = 0; i < n_sv; ++i) { s += static_cast<ld>(alpha[i]) * kernel_function(SV + i * d, P + pt * d, d, p); }
        out[pt] = static_cast<double>(s - static_cast<ld>(rho));
    }
}

// ---- the reference's arithmetic on sampled rows ------------------------------------------------------------------------------------------
// Rows of Q~ v evaluated like the reference's CPU kernels do it: every kernel value from one sequential FMA chain in the real type
// (include/plssvm/detail/operators.hpp:117-126, 161-171; kernel_function_types.hpp:75-97), the row sum accumulated in the real type
// (svm_kernel.cpp:36-47).  The reference's own kernel cannot evaluate single rows of a 65 536-point problem in test time; this gives its
// rounding behaviour for the full-size, row-sampled comparisons (|reference arithmetic - exact| beside |repo - exact|).
template <typename T>
T plain_kernel_function(const T *x, const T *y, const std::size_t d, const params<T> &p) {
    T s{};
    if (p.kernel == 2) {
        for (std::size_t k = 0; k < d; ++k) {
            const T diff = x[k] - y[k];
            s = std::fma(diff, diff, s);
        }
        return std::exp(-p.gamma * s);
    }
    for (std::size_t k = 0; k < d; ++k) { s = std::fma(x[k], y[k], s); }
    return p.kernel == 0 ? s : std::pow(std::fma(p.gamma, s, p.coef0), static_cast<T>(p.degree));
}
template <typename T>
void plain_matvec(const int kernel, const T *X, const std::size_t N, const std::size_t d, const T *q, const T *v, const T QA_cost, const T cost_inv, const int degree, const T gamma,
                  const T coef0, const std::uint64_t *rows, const std::size_t n_rows, T *out) {
    const std::size_t n = N - 1;
    const params<T> p{ kernel, degree, gamma, coef0 };
    #pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t r = 0; r < n_rows; ++r) {
        const std::size_t i = static_cast<std::size_t>(rows[r]);
        T s{};
        for (std::size_t j = 0; j < n; ++j) {
            const T temp = plain_kernel_function(X + i * d, X + j * d, d, p) + QA_cost - q[i] - q[j];
            s += (i == j ? temp + cost_inv : temp) * v[j];
        }
        out[r] = s;
    }
}

}  // namespace

extern "C" {

#define EXACT_INSTANTIATE(SUF, T)                                                                                                                                                   \
    int exact_solve_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, const T *y, int degree, T gamma, T coef0, T cost, T eps, std::uint64_t max_iter, double *alpha,    \
                          double *rho, std::uint64_t *iters, double *trace) {                                                                                                       \
        return exact_solve<T>(kernel, X, N, d, y, degree, gamma, coef0, cost, eps, max_iter, alpha, rho, iters, trace);                                                             \
    }                                                                                                                                                                               \
    void exact_matvec_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, const T *q, const T *v, T QA_cost, T cost_inv, int degree, T gamma, T coef0,                     \
                            const std::uint64_t *rows, std::size_t n_rows, double *out) {                                                                                          \
        exact_matvec<T>(kernel, X, N, d, q, v, QA_cost, cost_inv, degree, gamma, coef0, rows, n_rows, out);                                                                         \
    }                                                                                                                                                                               \
    void plain_matvec_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, const T *q, const T *v, T QA_cost, T cost_inv, int degree, T gamma, T coef0,                     \
                            const std::uint64_t *rows, std::size_t n_rows, T *out) {                                                                                               \
        plain_matvec<T>(kernel, X, N, d, q, v, QA_cost, cost_inv, degree, gamma, coef0, rows, n_rows, out);                                                                         \
    }                                                                                                                                                                               \
    void exact_q_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, int degree, T gamma, T coef0, double *q_out) { exact_q<T>(kernel, X, N, d, degree, gamma, coef0, q_out); } \
    void exact_predict_##SUF(int kernel, const T *SV, std::size_t n_sv, std::size_t d, const T *alpha, T rho, const T *P, std::size_t m, int degree, T gamma, T coef0,             \
                             double *out) {                                                                                                                                         \
        exact_predict<T>(kernel, SV, n_sv, d, alpha, rho, P, m, degree, gamma, coef0, out);                                                                                         \
    }

EXACT_INSTANTIATE(f32, float)
EXACT_INSTANTIATE(f64, double)

}  // extern "C"
