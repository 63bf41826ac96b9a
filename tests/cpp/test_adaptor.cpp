// Exercises include/plssvm_b200/csvm.hpp the way the reference's generic backend tests exercise a backend
// (tests/backends/generic_csvm_tests.hpp:99-137 solve_system_of_linear_equations_trivial, :149-195 predict_values).
// Without a CUDA device the constructor must throw backend_exception (reference: CUDA/csvm.cu:71-73) — exit code 3.
#include "plssvm_b200/csvm.hpp"

#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

template <typename T>
int run(const plssvm::b200::csvm &svm, const plssvm::b200::kernel_function_type kernel) {
    using namespace plssvm::b200;
    int failures = 0;
    parameter<T> params;
    params.kernel_type = kernel;
    params.cost = T{ 2.0 };
    params.degree = 1;
    params.gamma = T{ 1.0 };
    params.coef0 = T{ 0.0 };
    const T s = std::sqrt(T{ 1.0 } - T{ 1 } / params.cost);
    const std::vector<std::vector<T>> A = { { s, 0, 0, 0 }, { 0, s, 0, 0 }, { 0, 0, s, 0 }, { 0, 0, 0, s } };
    const std::vector<T> rhs{ 1, -1, 1, -1 };
    const auto [x, rho] = svm.solve_system_of_linear_equations(params, A, rhs, T{ 0.00001 }, 4ull);
    for (std::size_t i = 0; i < rhs.size(); ++i) {
        if (std::abs(x[i] - rhs[i]) > 256 * std::numeric_limits<T>::epsilon()) {
            std::printf("solve: x[%zu] = %g, expected %g\n", i, static_cast<double>(x[i]), static_cast<double>(rhs[i]));
            ++failures;
        }
    }
    if (std::abs(rho) > 8 * std::numeric_limits<T>::epsilon()) {
        std::printf("solve: rho = %g, expected 0\n", static_cast<double>(rho));
        ++failures;
    }

    const std::vector<std::vector<T>> sv = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
    const std::vector<T> weights{ 1, -1, 1, -1 };
    std::vector<T> w{};
    const std::vector<std::vector<T>> data{ { 1, 1, 1, 1 }, { 1, -1, 1, -1 } };
    const std::vector<T> vals = svm.predict_values(params, sv, weights, T{ 0 }, w, data);
    if (vals.size() != 2 || std::abs(vals[0]) > 64 * std::numeric_limits<T>::epsilon() || std::abs(vals[1] - T{ 4 }) > 64 * std::numeric_limits<T>::epsilon()) {
        std::printf("predict_values: got {%g, %g}, expected {0, 4}\n", static_cast<double>(vals[0]), static_cast<double>(vals[1]));
        ++failures;
    }
    if (kernel == kernel_function_type::linear) {
        if (w.size() != 4 || w[0] != 1 || w[1] != -1 || w[2] != 1 || w[3] != -1) {
            std::printf("predict_values: w not filled with the weights\n");
            ++failures;
        }
    } else if (!w.empty()) {
        std::printf("predict_values: w must stay empty for non-linear kernels\n");
        ++failures;
    }
    // argument errors surface as backend_exception (the reference asserts: gpu_csvm.hpp:484-489)
    try {
        (void) svm.solve_system_of_linear_equations(params, A, std::vector<T>{ 1, -1 }, T{ 0.1 }, 4ull);
        std::printf("size mismatch not detected\n");
        ++failures;
    } catch (const backend_exception &) {}
    try {  // ragged rows ("All data points must have the same number of features!", gpu_csvm.hpp:486)
        std::vector<std::vector<T>> ragged = A;
        ragged[2].pop_back();
        (void) svm.solve_system_of_linear_equations(params, ragged, rhs, T{ 0.1 }, 4ull);
        std::printf("ragged rows not detected\n");
        ++failures;
    } catch (const backend_exception &) {}
    try {  // empty data ("The data must not be empty!", gpu_csvm.hpp:484)
        (void) svm.solve_system_of_linear_equations(params, std::vector<std::vector<T>>{}, std::vector<T>{}, T{ 0.1 }, 4ull);
        std::printf("empty data not detected\n");
        ++failures;
    } catch (const backend_exception &) {}
    try {  // max_iter == 0 (gpu_csvm.hpp:489)
        (void) svm.solve_system_of_linear_equations(params, A, rhs, T{ 0.1 }, 0ull);
        std::printf("max_iter = 0 not detected\n");
        ++failures;
    } catch (const backend_exception &) {}
    try {
        (void) svm.solve_system_of_linear_equations(params, A, rhs, T{ 0 }, 4ull);
        std::printf("eps = 0 not detected\n");
        ++failures;
    } catch (const backend_exception &) {}
    return failures;
}

int main() {
    using plssvm::b200::kernel_function_type;
    try {
        const plssvm::b200::csvm svm{ 0 };
        int failures = 0;
        for (const kernel_function_type k : { kernel_function_type::linear, kernel_function_type::polynomial }) {
            failures += run<double>(svm, k);
            failures += run<float>(svm, k);
        }
        std::printf(failures == 0 ? "adaptor tests passed\n" : "adaptor tests FAILED (%d)\n", failures);
        return failures == 0 ? 0 : 1;
    } catch (const plssvm::b200::backend_exception &e) {
        std::printf("backend_exception (code %d): %s\n", e.code(), e.what());
        return 3;
    }
}
