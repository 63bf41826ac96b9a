// C bridge over the UNMODIFIED reference library (compiled in place from /root/reference by integration/Makefile) for tests:
//   backend 0 = the reference's own plssvm::openmp::csvm, backend 1 = plssvm::b200x::csvm (integration/b200_csvm.hpp);
// both are driven through the reference's public API — data_set, csvm::fit / predict / score, model::save / load — so the
// b200 backend is exercised exactly the way plssvm-train / plssvm-predict use a backend (main_train.cpp:42-51,
// main_predict.cpp:47-58), and the reference's real CG driver (OpenMP/csvm.cpp:71-183) is available as an oracle.
#include "b200_csvm.hpp"

#include "plssvm/backends/OpenMP/csvm.hpp"
#include "plssvm/csvm.hpp"
#include "plssvm/data_set.hpp"
#include "plssvm/detail/logger.hpp"
#include "plssvm/kernel_function_types.hpp"
#include "plssvm/model.hpp"
#include "plssvm/parameter.hpp"

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

template <typename T>
std::vector<std::vector<T>> rows(const T *flat, const std::size_t n, const std::size_t d) {
    std::vector<std::vector<T>> m(n);
    for (std::size_t i = 0; i < n; ++i) { m[i].assign(flat + i * d, flat + (i + 1) * d); }
    return m;
}

plssvm::parameter make_params(const int kernel, const int degree, const double gamma, const double coef0, const double cost) {
    plssvm::parameter p;
    p.kernel_type = static_cast<plssvm::kernel_function_type>(kernel);
    p.degree = degree;
    if (gamma > 0.0) { p.gamma = gamma; }
    p.coef0 = coef0;
    p.cost = cost;
    return p;
}

std::unique_ptr<plssvm::csvm> make_backend(const int backend, const plssvm::parameter &params) {
    if (backend == 0) { return std::make_unique<plssvm::openmp::csvm>(params); }
    return std::make_unique<plssvm::b200x::csvm>(params);
}

// exposes the protected virtuals of the reference's OpenMP backend, like the reference's own test mocks do
// (tests/backends/OpenMP/mock_openmp_csvm.hpp)
struct open_openmp_csvm : plssvm::openmp::csvm {
    using plssvm::openmp::csvm::csvm;
    using plssvm::openmp::csvm::predict_values;
    using plssvm::openmp::csvm::solve_system_of_linear_equations;
};

template <typename F>
int guarded(F &&f) {
    try {
        plssvm::verbosity = plssvm::verbosity_level::quiet;
        f();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    }
}

}  // namespace

extern "C" {

const char *refb_last_error() { return g_error.c_str(); }

// csvm::fit through the reference API.  labels: any two distinct ints.  Outputs: alpha[N] (model weights), rho, and optionally a LIBSVM model file.
int refb_fit_f64(const int backend, const double *X, const std::size_t N, const std::size_t d, const int *labels, const int kernel, const int degree, const double gamma,
                 const double coef0, const double cost, const double eps, const unsigned long long max_iter, double *alpha_out, double *rho_out, const char *model_path) {
    return guarded([&] {
        const plssvm::data_set<double, int> data{ rows(X, N, d), std::vector<int>(labels, labels + N) };
        const auto svm = make_backend(backend, make_params(kernel, degree, gamma, coef0, cost));
        const plssvm::model<double, int> model = svm->fit(data, plssvm::epsilon = eps, plssvm::max_iter = max_iter);
        std::copy(model.weights().begin(), model.weights().end(), alpha_out);
        *rho_out = model.rho();
        if (model_path != nullptr && model_path[0] != '\0') { model.save(model_path); }
    });
}

// the same, timing csvm::fit alone (seconds_out[0]) and the construction of the reference's data_set before it (seconds_out[1]): bench.py's `e2e_csvm`
int refb_fit_timed_f64(const int backend, const double *X, const std::size_t N, const std::size_t d, const int *labels, const int kernel, const int degree, const double gamma,
                       const double coef0, const double cost, const double eps, const unsigned long long max_iter, double *alpha_out, double *rho_out, double *seconds_out) {
    return guarded([&] {
        const auto t0 = std::chrono::steady_clock::now();
        const plssvm::data_set<double, int> data{ rows(X, N, d), std::vector<int>(labels, labels + N) };
        const auto svm = make_backend(backend, make_params(kernel, degree, gamma, coef0, cost));
        const auto t1 = std::chrono::steady_clock::now();
        const plssvm::model<double, int> model = svm->fit(data, plssvm::epsilon = eps, plssvm::max_iter = max_iter);
        const auto t2 = std::chrono::steady_clock::now();
        std::copy(model.weights().begin(), model.weights().end(), alpha_out);
        *rho_out = model.rho();
        seconds_out[0] = std::chrono::duration<double>(t2 - t1).count();
        seconds_out[1] = std::chrono::duration<double>(t1 - t0).count();
    });
}

// csvm::predict + csvm::score through the reference API on a model file written by either backend (LIBSVM model format)
int refb_predict_f64(const int backend, const char *model_path, const double *P, const std::size_t m, const std::size_t d, const int *true_labels, int *labels_out, double *score_out) {
    return guarded([&] {
        const plssvm::model<double, int> model{ std::string{ model_path } };
        const auto svm = make_backend(backend, plssvm::parameter{});
        if (true_labels != nullptr) {
            const plssvm::data_set<double, int> data{ rows(P, m, d), std::vector<int>(true_labels, true_labels + m) };
            const std::vector<int> pred = svm->predict(model, data);
            std::copy(pred.begin(), pred.end(), labels_out);
            if (score_out != nullptr) { *score_out = svm->score(model, data); }
        } else {
            const plssvm::data_set<double, int> data{ rows(P, m, d) };
            const std::vector<int> pred = svm->predict(model, data);
            std::copy(pred.begin(), pred.end(), labels_out);
        }
    });
}

// the reference's REAL CG driver: openmp::csvm::solve_system_of_linear_equations (OpenMP/csvm.cpp:71-183), y = +-1 values
int refb_openmp_solve_f64(const double *X, const std::size_t N, const std::size_t d, const double *y, const int kernel, const int degree, const double gamma, const double coef0,
                          const double cost, const double eps, const unsigned long long max_iter, double *alpha_out, double *rho_out) {
    return guarded([&] {
        const plssvm::parameter base = make_params(kernel, degree, gamma > 0.0 ? gamma : 1.0 / static_cast<double>(d), coef0, cost);
        const open_openmp_csvm svm{ base };
        const plssvm::detail::parameter<double> params{ base };
        const auto [alpha, rho] = svm.solve_system_of_linear_equations(params, rows(X, N, d), std::vector<double>(y, y + N), eps, max_iter);
        std::copy(alpha.begin(), alpha.end(), alpha_out);
        *rho_out = rho;
    });
}

int refb_openmp_solve_f32(const float *X, const std::size_t N, const std::size_t d, const float *y, const int kernel, const int degree, const float gamma, const float coef0,
                          const float cost, const float eps, const unsigned long long max_iter, float *alpha_out, float *rho_out) {
    return guarded([&] {
        const plssvm::parameter base = make_params(kernel, degree, gamma > 0.0f ? gamma : 1.0 / static_cast<double>(d), coef0, cost);
        const open_openmp_csvm svm{ base };
        const plssvm::detail::parameter<float> params{ base };
        const auto [alpha, rho] = svm.solve_system_of_linear_equations(params, rows(X, N, d), std::vector<float>(y, y + N), eps, max_iter);
        std::copy(alpha.begin(), alpha.end(), alpha_out);
        *rho_out = rho;
    });
}

// the reference's REAL predict path: openmp::csvm::predict_values (OpenMP/csvm.cpp:188-227)
int refb_openmp_predict_values_f64(const double *SV, const std::size_t n_sv, const std::size_t d, const double *alpha, const double rho, const double *P, const std::size_t m,
                                   const int kernel, const int degree, const double gamma, const double coef0, double *out) {
    return guarded([&] {
        const plssvm::parameter base = make_params(kernel, degree, gamma > 0.0 ? gamma : 1.0 / static_cast<double>(d), coef0, 1.0);
        const open_openmp_csvm svm{ base };
        const plssvm::detail::parameter<double> params{ base };
        std::vector<double> w;
        const std::vector<double> res = svm.predict_values(params, rows(SV, n_sv, d), std::vector<double>(alpha, alpha + n_sv), rho, w, rows(P, m, d));
        std::copy(res.begin(), res.end(), out);
    });
}

}  // extern "C"
