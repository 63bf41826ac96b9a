// TEST INFRASTRUCTURE ONLY — C ABI of the parity oracle (see lssvm_oracle.hpp).
// Built twice by oracle/Makefile:
//   default                      -> liboracle_port.so       (restated kernels)
//   -DORACLE_USE_REFERENCE=1     -> _ref/liboracle_ref.so   (links the reference's OpenMP kernel TUs, compiled in place)
#include "lssvm_oracle.hpp"

#include <cstring>
#include <omp.h>

#if defined(ORACLE_USE_REFERENCE)
    // the reference's own headers, from /root/reference/include (never copied into this repo)
    #include "plssvm/backends/OpenMP/q_kernel.hpp"
    #include "plssvm/backends/OpenMP/svm_kernel.hpp"
    #include "plssvm/kernel_function_types.hpp"

namespace oracle {
struct ref_kernels {
    static const char *kind() { return "reference"; }

    template <typename T>
    static T kernel_function(const std::vector<T> &xi, const std::vector<T> &xj, const params<T> &p) {
        using kt = plssvm::kernel_function_type;
        switch (p.kernel) {
            case k_linear:
                return plssvm::kernel_function<kt::linear>(xi, xj);
            case k_polynomial:
                return plssvm::kernel_function<kt::polynomial>(xi, xj, p.degree, p.gamma, p.coef0);
            default:
                return plssvm::kernel_function<kt::rbf>(xi, xj, p.gamma);
        }
    }
    template <typename T>
    static void q_kernel(std::vector<T> &q, const std::vector<std::vector<T>> &data, const params<T> &p) {
        switch (p.kernel) {
            case k_linear:
                plssvm::openmp::device_kernel_q_linear(q, data);
                break;
            case k_polynomial:
                plssvm::openmp::device_kernel_q_polynomial(q, data, p.degree, p.gamma, p.coef0);
                break;
            default:
                plssvm::openmp::device_kernel_q_rbf(q, data, p.gamma);
                break;
        }
    }
    template <typename T>
    static void svm_kernel(const std::vector<T> &q, std::vector<T> &ret, const std::vector<T> &d, const std::vector<std::vector<T>> &data, const T QA_cost, const T cost, const T add, const params<T> &p) {
        switch (p.kernel) {
            case k_linear:
                plssvm::openmp::device_kernel_linear(q, ret, d, data, QA_cost, cost, add);
                break;
            case k_polynomial:
                plssvm::openmp::device_kernel_polynomial(q, ret, d, data, QA_cost, cost, add, p.degree, p.gamma, p.coef0);
                break;
            default:
                plssvm::openmp::device_kernel_rbf(q, ret, d, data, QA_cost, cost, add, p.gamma);
                break;
        }
    }
};
using provider = ref_kernels;
}  // namespace oracle
#else
namespace oracle {
using provider = port_kernels;
}
#endif

namespace {

template <typename T>
T c_kernel_function(int kernel, const T *x, const T *y, std::size_t d, int degree, T gamma, T coef0) {
    const oracle::params<T> p{ kernel, degree, gamma, coef0, T{ 1 } };
    return oracle::provider::kernel_function(std::vector<T>(x, x + d), std::vector<T>(y, y + d), p);
}

template <typename T>
void c_q(int kernel, const T *X, std::size_t N, std::size_t d, int degree, T gamma, T coef0, T *q) {
    const oracle::params<T> p{ kernel, degree, gamma, coef0, T{ 1 } };
    const auto data = oracle::to_rows(X, N, d);
    std::vector<T> qv(N - 1);
    oracle::provider::q_kernel(qv, data, p);
    std::copy(qv.begin(), qv.end(), q);
}

template <typename T>
void c_matvec(int kernel, const T *X, std::size_t N, std::size_t d, const T *q, const T *v, T *ret, T QA_cost, T cost, T add, int degree, T gamma, T coef0) {
    const oracle::params<T> p{ kernel, degree, gamma, coef0, T{ 1 } / cost };
    const auto data = oracle::to_rows(X, N, d);
    const std::size_t n = N - 1;
    std::vector<T> rv(ret, ret + n);
    oracle::provider::svm_kernel(std::vector<T>(q, q + n), rv, std::vector<T>(v, v + n), data, QA_cost, cost, add, p);
    std::copy(rv.begin(), rv.end(), ret);
}

template <typename T>
int c_solve(int kernel, const T *X, std::size_t N, std::size_t d, const T *y, int degree, T gamma, T coef0, T cost, T eps, std::uint64_t max_iter, T *alpha, T *rho, std::uint64_t *iters, T *delta_out, T *delta_trace) {
    if (N < 2 || d == 0 || !(eps > T{ 0 }) || max_iter == 0) {
        return 1;
    }
    const oracle::params<T> p{ kernel, degree, gamma, coef0, cost };
    const auto data = oracle::to_rows(X, N, d);
    auto res = oracle::solve<oracle::provider>(p, data, std::vector<T>(y, y + N), eps, max_iter, delta_trace);
    std::copy(res.alpha.begin(), res.alpha.end(), alpha);
    *rho = res.rho;
    if (iters != nullptr) {
        *iters = res.iterations;
    }
    if (delta_out != nullptr) {
        delta_out[0] = res.delta;
        delta_out[1] = res.delta0;
    }
    return 0;
}

template <typename T>
void c_w(const T *SV, std::size_t n_sv, std::size_t d, const T *alpha, T *w) {
    const auto w_vec = oracle::calculate_w(oracle::to_rows(SV, n_sv, d), std::vector<T>(alpha, alpha + n_sv));
    std::copy(w_vec.begin(), w_vec.end(), w);
}

template <typename T>
void c_predict(int kernel, const T *SV, std::size_t n_sv, std::size_t d, const T *alpha, T rho, T *w_inout, int *w_valid, const T *P, std::size_t m, int degree, T gamma, T coef0, T *out) {
    const oracle::params<T> p{ kernel, degree, gamma, coef0, T{ 1 } };
    std::vector<T> w;
    if (w_valid != nullptr && *w_valid != 0) {
        w.assign(w_inout, w_inout + d);
    }
    const auto res = oracle::predict_values<oracle::provider>(p, oracle::to_rows(SV, n_sv, d), std::vector<T>(alpha, alpha + n_sv), rho, w, oracle::to_rows(P, m, d));
    std::copy(res.begin(), res.end(), out);
    if (!w.empty() && w_inout != nullptr) {
        std::copy(w.begin(), w.end(), w_inout);
        if (w_valid != nullptr) {
            *w_valid = 1;
        }
    }
}

}  // namespace

extern "C" {

const char *oracle_kind(void) { return oracle::provider::kind(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }

#define ORACLE_INSTANTIATE(SUF, T)                                                                                                                                           \
    T oracle_kernel_function_##SUF(int kernel, const T *x, const T *y, std::size_t d, int degree, T gamma, T coef0) {                                                        \
        return c_kernel_function<T>(kernel, x, y, d, degree, gamma, coef0);                                                                                                  \
    }                                                                                                                                                                        \
    void oracle_q_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, int degree, T gamma, T coef0, T *q) {                                                          \
        c_q<T>(kernel, X, N, d, degree, gamma, coef0, q);                                                                                                                    \
    }                                                                                                                                                                        \
    void oracle_matvec_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, const T *q, const T *v, T *ret, T QA_cost, T cost, T add, int degree, T gamma, T coef0) { \
        c_matvec<T>(kernel, X, N, d, q, v, ret, QA_cost, cost, add, degree, gamma, coef0);                                                                                   \
    }                                                                                                                                                                        \
    int oracle_solve_##SUF(int kernel, const T *X, std::size_t N, std::size_t d, const T *y, int degree, T gamma, T coef0, T cost, T eps, std::uint64_t max_iter,            \
                           T *alpha, T *rho, std::uint64_t *iters, T *delta_out, T *delta_trace) {                                                                           \
        return c_solve<T>(kernel, X, N, d, y, degree, gamma, coef0, cost, eps, max_iter, alpha, rho, iters, delta_out, delta_trace);                                         \
    }                                                                                                                                                                        \
    void oracle_w_##SUF(const T *SV, std::size_t n_sv, std::size_t d, const T *alpha, T *w) {                                                                                \
        c_w<T>(SV, n_sv, d, alpha, w);                                                                                                                                       \
    }                                                                                                                                                                        \
    void oracle_predict_##SUF(int kernel, const T *SV, std::size_t n_sv, std::size_t d, const T *alpha, T rho, T *w_inout, int *w_valid, const T *P, std::size_t m,          \
                              int degree, T gamma, T coef0, T *out) {                                                                                                        \
        c_predict<T>(kernel, SV, n_sv, d, alpha, rho, w_inout, w_valid, P, m, degree, gamma, coef0, out);                                                                    \
    }

ORACLE_INSTANTIATE(f32, float)
ORACLE_INSTANTIATE(f64, double)

}  // extern "C"
