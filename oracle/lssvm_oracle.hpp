// TEST INFRASTRUCTURE ONLY — the parity oracle.  Nothing under plssvm_b200/ may include, link or call this.
//
// CPU restatement of the reference's LS-SVM hot path (PLSSVM v2.0.0, OpenMP backend).  Every function cites the
// reference file:line it follows.  The driver part (CG loop, predict loop, w) is shared between two builds:
//   * liboracle_port.so  — kernels restated here (struct port_kernels below);                 kind = "port"
//   * _ref/liboracle_ref.so — kernels are the reference's own OpenMP translation units
//     (src/plssvm/backends/OpenMP/{svm_kernel,q_kernel}.cpp + kernel_function_types.hpp) compiled in place from
//     /root/reference by oracle/Makefile;                                                       kind = "reference"
// Parity pinning: tests/test_oracle_golden.py checks both builds against the reference's known-answer tests
// (tests/backends/generic_csvm_tests.hpp:99-137,149-195,197-247) and against each other.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace oracle {

enum kernel_id : int { k_linear = 0,
                       k_polynomial = 1,
                       k_rbf = 2 };

template <typename T>
struct params {
    int kernel;
    int degree;
    T gamma;
    T coef0;
    T cost;
};

using matrix_f32 = std::vector<std::vector<float>>;
using matrix_f64 = std::vector<std::vector<double>>;

template <typename T>
std::vector<std::vector<T>> to_rows(const T *flat, const std::size_t rows, const std::size_t cols) {
    std::vector<std::vector<T>> m(rows);
    for (std::size_t r = 0; r < rows; ++r) {
        m[r].assign(flat + r * cols, flat + (r + 1) * cols);
    }
    return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// vector algebra: include/plssvm/detail/operators.hpp
// ---------------------------------------------------------------------------------------------------------------------

// operators.hpp:117-126 — strictly sequential FMA chain
template <typename T>
T dot(const std::vector<T> &a, const std::vector<T> &b) {
    T val{};
    for (std::size_t i = 0; i < a.size(); ++i) {
        val = std::fma(a[i], b[i], val);
    }
    return val;
}

// operators.hpp:143-151 — `omp simd reduction(+)`
template <typename T>
T sum(const std::vector<T> &v) {
    T val{};
    #pragma omp simd reduction(+ : val)
    for (std::size_t i = 0; i < v.size(); ++i) {
        val += v[i];
    }
    return val;
}

// operators.hpp:161-171
template <typename T>
T squared_euclidean_dist(const std::vector<T> &a, const std::vector<T> &b) {
    T val{};
    for (std::size_t i = 0; i < a.size(); ++i) {
        const T diff = a[i] - b[i];
        val = std::fma(diff, diff, val);
    }
    return val;
}

// operators.hpp:178-181 — sign(0) == -1
template <typename T>
constexpr T sign(const T x) {
    return x > T{ 0 } ? T{ +1 } : T{ -1 };
}

// ---------------------------------------------------------------------------------------------------------------------
// restated kernels ("port" provider)
// ---------------------------------------------------------------------------------------------------------------------

struct port_kernels {
    static const char *kind() { return "port"; }

    // include/plssvm/kernel_function_types.hpp:75-97
    template <typename T>
    static T kernel_function(const std::vector<T> &xi, const std::vector<T> &xj, const params<T> &p) {
        switch (p.kernel) {
            case k_linear:
                return dot(xi, xj);
            case k_polynomial:
                return std::pow(std::fma(p.gamma, dot(xi, xj), p.coef0), static_cast<T>(p.degree));
            default:
                return std::exp(-p.gamma * squared_euclidean_dist(xi, xj));
        }
    }

    // src/plssvm/backends/OpenMP/q_kernel.cpp:18-52 — q_i = k(x_i, x_last), i < N-1
    template <typename T>
    static void q_kernel(std::vector<T> &q, const std::vector<std::vector<T>> &data, const params<T> &p) {
        #pragma omp parallel for
        for (std::size_t i = 0; i < data.size() - 1; ++i) {
            q[i] = kernel_function(data[i], data.back(), p);
        }
    }

    // src/plssvm/backends/OpenMP/svm_kernel.cpp:22-55 — ret += add * Q~ * d over the lower triangle, mirrored, 64x64 blocks
    template <typename T>
    static void svm_kernel(const std::vector<T> &q, std::vector<T> &ret, const std::vector<T> &d, const std::vector<std::vector<T>> &data, const T QA_cost, const T cost, const T add, const params<T> &p) {
        constexpr long BLOCK = 64;  // include/plssvm/constants.hpp:39 OPENMP_BLOCK_SIZE
        const long dept = static_cast<long>(d.size());
        #pragma omp parallel for collapse(2) schedule(dynamic)
        for (long i = 0; i < dept; i += BLOCK) {
            for (long j = 0; j < dept; j += BLOCK) {
                for (long ii = 0; ii < BLOCK && ii + i < dept; ++ii) {
                    T ret_iii = 0.0;
                    for (long jj = 0; jj < BLOCK && jj + j < dept; ++jj) {
                        if (ii + i >= jj + j) {
                            const T temp = (kernel_function(data[ii + i], data[jj + j], p) + QA_cost - q[ii + i] - q[jj + j]) * add;
                            if (ii + i == jj + j) {
                                ret_iii += (temp + cost * add) * d[ii + i];
                            } else {
                                ret_iii += temp * d[jj + j];
                                #pragma omp atomic
                                ret[jj + j] += temp * d[ii + i];
                            }
                        }
                    }
                    #pragma omp atomic
                    ret[ii + i] += ret_iii;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// drivers (shared by both builds)
// ---------------------------------------------------------------------------------------------------------------------

template <typename T>
struct solve_result {
    std::vector<T> alpha;  // length N
    T rho;
    std::uint64_t iterations;  // min(iter + 1, max_iter)
    T delta;                   // final residual r.r
    T delta0;
};

// src/plssvm/backends/OpenMP/csvm.cpp:71-183 (identical algorithm: include/plssvm/backends/gpu_csvm.hpp:477-654)
template <typename K, typename T>
solve_result<T> solve(const params<T> &p, const std::vector<std::vector<T>> &A, std::vector<T> b, const T eps, const std::uint64_t max_iter, T *delta_trace /* optional, max_iter + 1 entries */) {
    // q vector (csvm.cpp:83, 232-251)
    std::vector<T> q(A.size() - 1);
    K::q_kernel(q, A, p);
    // QA_cost (csvm.cpp:86)
    const T QA_cost = K::kernel_function(A.back(), A.back(), p) + T{ 1.0 } / p.cost;
    // update b (csvm.cpp:89-91)
    const T b_back_value = b.back();
    b.pop_back();
    for (T &v : b) {
        v -= b_back_value;
    }
    // CG (csvm.cpp:95-160)
    std::vector<T> alpha(b.size(), 1.0);
    const std::size_t dept = b.size();
    std::vector<T> r(b);
    K::svm_kernel(q, r, alpha, A, QA_cost, T{ 1 } / p.cost, T{ -1.0 }, p);
    T delta = dot(r, r);
    const T delta0 = delta;
    std::vector<T> Ad(dept);
    std::vector<T> d(r);
    if (delta_trace != nullptr) {
        delta_trace[0] = delta0;
    }

    std::uint64_t iter = 0;
    for (; iter < max_iter; ++iter) {
        std::fill(Ad.begin(), Ad.end(), T{ 0.0 });
        K::svm_kernel(q, Ad, d, A, QA_cost, T{ 1 } / p.cost, T{ 1.0 }, p);
        const T alpha_cd = delta / dot(d, Ad);
        #pragma omp simd
        for (std::size_t i = 0; i < dept; ++i) {
            alpha[i] += alpha_cd * d[i];
        }
        if (iter % 50 == 49) {
            r = b;
            K::svm_kernel(q, r, alpha, A, QA_cost, T{ 1 } / p.cost, T{ -1.0 }, p);
        } else {
            #pragma omp simd
            for (std::size_t i = 0; i < dept; ++i) {
                r[i] -= alpha_cd * Ad[i];
            }
        }
        const T delta_old = delta;
        delta = dot(r, r);
        if (delta_trace != nullptr) {
            delta_trace[iter + 1] = delta;
        }
        if (delta <= eps * eps * delta0) {
            break;
        }
        const T beta = delta / delta_old;
        #pragma omp simd
        for (std::size_t i = 0; i < dept; ++i) {
            d[i] = beta * d[i] + r[i];
        }
    }
    // bias (csvm.cpp:176-180)
    const T bias = b_back_value + QA_cost * sum(alpha) - dot(q, alpha);
    alpha.push_back(-sum(alpha));
    return solve_result<T>{ std::move(alpha), -bias, std::min<std::uint64_t>(iter + 1, max_iter), delta, delta0 };
}

// src/plssvm/backends/OpenMP/csvm.cpp:255-280
template <typename T>
std::vector<T> calculate_w(const std::vector<std::vector<T>> &sv, const std::vector<T> &alpha) {
    const std::size_t num_data_points = sv.size();
    const std::size_t num_features = sv.front().size();
    std::vector<T> w(num_features, T{ 0.0 });
    #pragma omp parallel for
    for (std::size_t f = 0; f < num_features; ++f) {
        T temp{ 0.0 };
        #pragma omp simd reduction(+ : temp)
        for (std::size_t i = 0; i < num_data_points; ++i) {
            temp = std::fma(alpha[i], sv[i][f], temp);
        }
        w[f] = temp;
    }
    return w;
}

// src/plssvm/backends/OpenMP/csvm.cpp:188-227
template <typename K, typename T>
std::vector<T> predict_values(const params<T> &p, const std::vector<std::vector<T>> &sv, const std::vector<T> &alpha, const T rho, std::vector<T> &w, const std::vector<std::vector<T>> &points) {
    std::vector<T> out(points.size(), -rho);
    if (p.kernel == k_linear && w.empty()) {
        w = calculate_w(sv, alpha);
    }
    #pragma omp parallel for
    for (std::size_t pt = 0; pt < points.size(); ++pt) {
        if (p.kernel == k_linear) {
            out[pt] += dot(w, points[pt]);
        } else {
            T temp{ 0.0 };
            #pragma omp simd reduction(+ : temp)
            for (std::size_t i = 0; i < sv.size(); ++i) {
                temp += alpha[i] * K::kernel_function(sv[i], points[pt], p);
            }
            out[pt] += temp;
        }
    }
    return out;
}

}  // namespace oracle
