/**
 * plssvm::b200::csvm — C++ host adaptor over the C ABI (include/plssvm_b200.h), shaped like a backend of the reference.
 *
 * A backend of PLSSVM v2.0.0 is a class deriving `plssvm::csvm` (include/plssvm/csvm.hpp:50) that overrides four protected
 * const virtuals (csvm.hpp:188-208):
 *
 *   std::pair<std::vector<T>, T> solve_system_of_linear_equations(const detail::parameter<T>&, const std::vector<std::vector<T>>& A,
 *                                                                 std::vector<T> b, T eps, unsigned long long max_iter) const;
 *   std::vector<T> predict_values(const detail::parameter<T>&, const std::vector<std::vector<T>>& support_vectors, const std::vector<T>& alpha,
 *                                 T rho, std::vector<T>& w, const std::vector<std::vector<T>>& predict_points) const;            (T = float, double)
 *
 * This header provides exactly those member functions with the same argument meaning, ownership and error behaviour
 * (SURVEY.md §8b), handing the rows of the reference's `std::vector<std::vector<T>>` to the C ABI as row pointers (the library
 * stages them through a pinned ring; no flat host copy).  It is self-contained (STL only) so that it compiles without the reference's third-party headers (igor, fmt,
 * fast_float are not available offline); INTEGRATION.md shows the ~40-line `class csvm : public ::plssvm::csvm` that
 * forwards the reference's virtuals to this class inside the reference tree, and the enum / factory registration.
 */
#ifndef PLSSVM_B200_CSVM_HPP_
#define PLSSVM_B200_CSVM_HPP_

#include "../plssvm_b200.h"

#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace plssvm::b200 {

/// ~ plssvm::cuda::backend_exception (include/plssvm/backends/CUDA/exceptions.hpp:26-34)
class backend_exception : public std::runtime_error {
  public:
    explicit backend_exception(const std::string &msg, const int code = PLSSVM_B200_ERR_CUDA) : std::runtime_error{ msg }, code_{ code } {}
    [[nodiscard]] int code() const noexcept { return code_; }

  private:
    int code_;
};

/// ~ plssvm::kernel_function_type (include/plssvm/kernel_function_types.hpp:31-38), same enumerator values
enum class kernel_function_type { linear = PLSSVM_B200_KERNEL_LINEAR, polynomial = PLSSVM_B200_KERNEL_POLYNOMIAL, rbf = PLSSVM_B200_KERNEL_RBF };

/// ~ plssvm::detail::parameter<T> (include/plssvm/parameter.hpp:105-266) reduced to the fields that reach the virtuals; defaults parameter.hpp:157-165
template <typename T>
struct parameter {
    kernel_function_type kernel_type{ kernel_function_type::linear };
    int degree{ 3 };
    T gamma{ 0 };  ///< 0 = "not set": 1 / #features is used, as csvm::fit does (csvm.hpp:304-307)
    T coef0{ 0 };
    T cost{ 1 };
};

namespace detail {

inline void check(const int rc) {
    if (rc != PLSSVM_B200_OK) {
        throw backend_exception{ plssvm_b200_last_error(), rc };
    }
}

/// Row pointers of a `std::vector<std::vector<T>>` for the *_rows entry points: the library stages the rows straight into a pinned ring
/// (no flat host copy — at 1 M x 4096 doubles that copy alone would be 34 GB).  Checks the shape like the reference's asserts (gpu_csvm.hpp:484-489).
template <typename T>
[[nodiscard]] std::vector<const T *> row_pointers(const std::vector<std::vector<T>> &rows, const char *what) {
    if (rows.empty()) {
        throw backend_exception{ std::string{ "The " } + what + " must not be empty!", PLSSVM_B200_ERR_INVALID };
    }
    const std::size_t d = rows.front().size();
    if (d == 0) {
        throw backend_exception{ std::string{ "The " } + what + " must contain at least one feature!", PLSSVM_B200_ERR_INVALID };
    }
    std::vector<const T *> ptrs(rows.size());
    for (std::size_t i = 0; i < rows.size(); ++i) {
        if (rows[i].size() != d) {
            throw backend_exception{ std::string{ "All " } + what + " must have the same number of features!", PLSSVM_B200_ERR_INVALID };
        }
        ptrs[i] = rows[i].data();
    }
    return ptrs;
}

}  // namespace detail

/// The b200 backend.  Move-only like every reference backend (csvm.hpp:69-83); one instance owns the context of the GPUs it drives.
class csvm {
  public:
    /// tag for an empty (context-less) object that is move-assigned later
    struct deferred {};
    /// tag: every visible device, like the reference's CUDA backend (CUDA/csvm.cu:66-68)
    struct all_devices {};
    explicit csvm(deferred) noexcept {}
    explicit csvm(const int device = 0) {
        detail::check(plssvm_b200_create(&device, 1, &ctx_));  // throws "…no CUDA devices were found!" like csvm.cu:71-73
    }
    explicit csvm(all_devices) {
        detail::check(plssvm_b200_create(nullptr, 0, &ctx_));
    }
    /// a device group: data sets replicated, matvec tiles and predict points sharded over the devices
    explicit csvm(const std::vector<int> &devices) {
        detail::check(plssvm_b200_create(devices.data(), static_cast<int>(devices.size()), &ctx_));
    }
    csvm(const csvm &) = delete;
    csvm &operator=(const csvm &) = delete;
    csvm(csvm &&other) noexcept : ctx_{ std::exchange(other.ctx_, nullptr) } {}
    csvm &operator=(csvm &&other) noexcept {
        if (this != &other) {
            plssvm_b200_destroy(ctx_);
            ctx_ = std::exchange(other.ctx_, nullptr);
        }
        return *this;
    }
    ~csvm() { plssvm_b200_destroy(ctx_); }

    [[nodiscard]] int num_devices() const {
        int n = 0;
        detail::check(plssvm_b200_num_devices(ctx_, &n));
        return n;
    }
    /// timings and CG statistics of the last call (the reference's `cg.*` tracker entries, gpu_csvm.hpp:637-646)
    [[nodiscard]] plssvm_b200_timings last_stats() const {
        plssvm_b200_timings t{};
        detail::check(plssvm_b200_get_timings(ctx_, &t));
        return t;
    }
    /// residual history r.r of the last solve: entry k = after k iterations (gpu_csvm.hpp:569-571 logs it per iteration)
    [[nodiscard]] std::vector<double> last_trace() const {
        std::vector<double> tr(4097);
        std::size_t count = 0;
        detail::check(plssvm_b200_last_trace(ctx_, tr.data(), tr.size(), &count));
        tr.resize(count);
        return tr;
    }

    /// csvm::solve_system_of_linear_equations (csvm.hpp:188-192; gpu_csvm.hpp:477-654): returns (alpha[N], rho)
    template <typename T>
    [[nodiscard]] std::pair<std::vector<T>, T> solve_system_of_linear_equations(const parameter<T> &params, const std::vector<std::vector<T>> &A, std::vector<T> b,
                                                                                const T eps, const unsigned long long max_iter, unsigned long long *iterations = nullptr) const {
        static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "real_type must be float or double");
        const std::vector<const T *> rows = detail::row_pointers(A, "data points");
        if (A.size() != b.size()) {
            throw backend_exception{ "The number of data points in the matrix A (" + std::to_string(A.size()) + ") and the values in the right hand side vector (" +
                                         std::to_string(b.size()) + ") must be the same!",
                                     PLSSVM_B200_ERR_INVALID };
        }
        const std::size_t N = A.size(), d = A.front().size();
        const T gamma = params.gamma > T{ 0 } ? params.gamma : T{ 1 } / static_cast<T>(d);
        std::vector<T> alpha(N);
        T rho{};
        uint64_t iters = 0;
        if constexpr (std::is_same_v<T, double>) {
            detail::check(plssvm_b200_solve_rows_f64(ctx_, rows.data(), N, d, b.data(), static_cast<int>(params.kernel_type), params.degree, gamma, params.coef0, params.cost, eps,
                                                max_iter, alpha.data(), &rho, &iters, nullptr));
        } else {
            detail::check(plssvm_b200_solve_rows_f32(ctx_, rows.data(), N, d, b.data(), static_cast<int>(params.kernel_type), params.degree, gamma, params.coef0, params.cost, eps,
                                                max_iter, alpha.data(), &rho, &iters, nullptr));
        }
        if (iterations != nullptr) {
            *iterations = iters;
        }
        return std::make_pair(std::move(alpha), rho);
    }

    /// csvm::predict_values (csvm.hpp:204-208; gpu_csvm.hpp:656-730): `w` is the caller-owned in/out cache, filled iff the kernel is linear and it is empty
    template <typename T>
    [[nodiscard]] std::vector<T> predict_values(const parameter<T> &params, const std::vector<std::vector<T>> &support_vectors, const std::vector<T> &alpha, const T rho,
                                                std::vector<T> &w, const std::vector<std::vector<T>> &predict_points) const {
        static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "real_type must be float or double");
        const std::vector<const T *> sv = detail::row_pointers(support_vectors, "support vectors");
        const std::vector<const T *> pts = detail::row_pointers(predict_points, "data points to predict");
        const std::size_t n_sv = support_vectors.size(), d = support_vectors.front().size(), m = predict_points.size();
        if (alpha.size() != n_sv) {
            throw backend_exception{ "The number of support vectors (" + std::to_string(n_sv) + ") and number of weights (" + std::to_string(alpha.size()) + ") must be the same!",
                                     PLSSVM_B200_ERR_INVALID };
        }
        if (predict_points.front().size() != d) {
            throw backend_exception{ "The number of features in the support vectors (" + std::to_string(d) + ") must be the same as in the data points to predict (" +
                                         std::to_string(predict_points.front().size()) + ")!",
                                     PLSSVM_B200_ERR_INVALID };
        }
        if (!w.empty() && w.size() != d) {
            throw backend_exception{ "Either w must be empty or contain exactly the same number of values (" + std::to_string(w.size()) + ") as features are present (" +
                                         std::to_string(d) + ")!",
                                     PLSSVM_B200_ERR_INVALID };
        }
        const T gamma = params.gamma > T{ 0 } ? params.gamma : T{ 1 } / static_cast<T>(d);
        std::vector<T> out(m);
        std::vector<T> w_buf(d, T{ 0 });
        int w_valid = 0;
        if (!w.empty()) {
            w_buf = w;
            w_valid = 1;
        }
        if constexpr (std::is_same_v<T, double>) {
            detail::check(plssvm_b200_predict_rows_f64(ctx_, sv.data(), n_sv, d, alpha.data(), rho, w_buf.data(), &w_valid, pts.data(), m, static_cast<int>(params.kernel_type),
                                                  params.degree, gamma, params.coef0, out.data()));
        } else {
            detail::check(plssvm_b200_predict_rows_f32(ctx_, sv.data(), n_sv, d, alpha.data(), rho, w_buf.data(), &w_valid, pts.data(), m, static_cast<int>(params.kernel_type),
                                                  params.degree, gamma, params.coef0, out.data()));
        }
        if (params.kernel_type == kernel_function_type::linear && w.empty() && w_valid != 0) {
            w = std::move(w_buf);  // gpu_csvm.hpp:696-698; stays empty for polynomial / rbf (generic_csvm_tests.hpp:188-194)
        }
        return out;
    }

    [[nodiscard]] plssvm_b200_ctx *native_handle() const noexcept { return ctx_; }

  private:
    plssvm_b200_ctx *ctx_{ nullptr };
};

}  // namespace plssvm::b200

#endif  // PLSSVM_B200_CSVM_HPP_
