/**
 * plssvm::b200x::csvm — the b200 backend as a REAL subclass of the unmodified reference base class `plssvm::csvm`
 * (include/plssvm/csvm.hpp:50-222 of PLSSVM v2.0.0): overrides the four protected virtuals (csvm.hpp:188-208) and forwards
 * them to the C ABI through the adaptor include/plssvm_b200/csvm.hpp.  With this class the reference's own `csvm::fit`,
 * `csvm::predict`, `csvm::score`, `data_set`, `model` and LIBSVM file I/O run on top of libplssvm_b200.so.
 * Compiled against the reference headers in place (integration/Makefile); constructor set and target handling mirror
 * cuda::csvm (include/plssvm/backends/CUDA/csvm.hpp:59-95; src/plssvm/backends/CUDA/csvm.cu:48-86).
 * The enum value `backend_type::b200` / the `make_csvm` case need a two-line edit of the reference;
 * until then the class is constructed directly.
 */
#ifndef PLSSVM_B200_INTEGRATION_CSVM_HPP_
#define PLSSVM_B200_INTEGRATION_CSVM_HPP_

#include "plssvm/csvm.hpp"                  // plssvm::csvm
#include "plssvm/detail/type_traits.hpp"    // PLSSVM_REQUIRES
#include "plssvm/exceptions/exceptions.hpp" // plssvm::exception
#include "plssvm/parameter.hpp"             // plssvm::parameter, plssvm::detail::parameter, has_only_parameter_named_args_v
#include "plssvm/target_platforms.hpp"      // plssvm::target_platform

#include "plssvm_b200/csvm.hpp"  // plssvm::b200::csvm (C-ABI adaptor)

#include "fmt/core.h"

#include <string>
#include <utility>
#include <vector>

namespace plssvm::b200x {

/// ~ plssvm::cuda::backend_exception (include/plssvm/backends/CUDA/exceptions.hpp:26-34): derives the reference's exception type
class backend_exception : public ::plssvm::exception {
  public:
    explicit backend_exception(const std::string &msg, source_location loc = source_location::current()) :
        ::plssvm::exception{ msg, "b200::backend_exception", loc } {}
};

class csvm : public ::plssvm::csvm {
  public:
    explicit csvm(parameter params = {}) : csvm{ plssvm::target_platform::automatic, params } {}
    explicit csvm(const target_platform target, parameter params = {}) : ::plssvm::csvm{ params } { this->init(target); }
    template <typename... Args, PLSSVM_REQUIRES(::plssvm::detail::has_only_parameter_named_args_v<Args...>)>
    explicit csvm(Args &&...named_args) : csvm{ plssvm::target_platform::automatic, std::forward<Args>(named_args)... } {}
    template <typename... Args, PLSSVM_REQUIRES(::plssvm::detail::has_only_parameter_named_args_v<Args...>)>
    explicit csvm(const target_platform target, Args &&...named_args) : ::plssvm::csvm{ std::forward<Args>(named_args)... } { this->init(target); }

    csvm(const csvm &) = delete;
    csvm(csvm &&) noexcept = default;
    csvm &operator=(const csvm &) = delete;
    csvm &operator=(csvm &&) noexcept = default;
    ~csvm() override = default;

  protected:
    [[nodiscard]] std::pair<std::vector<float>, float> solve_system_of_linear_equations(const ::plssvm::detail::parameter<float> &params, const std::vector<std::vector<float>> &A, std::vector<float> b, float eps, unsigned long long max_iter) const final { return this->solve_impl(params, A, std::move(b), eps, max_iter); }
    [[nodiscard]] std::pair<std::vector<double>, double> solve_system_of_linear_equations(const ::plssvm::detail::parameter<double> &params, const std::vector<std::vector<double>> &A, std::vector<double> b, double eps, unsigned long long max_iter) const final { return this->solve_impl(params, A, std::move(b), eps, max_iter); }
    [[nodiscard]] std::vector<float> predict_values(const ::plssvm::detail::parameter<float> &params, const std::vector<std::vector<float>> &support_vectors, const std::vector<float> &alpha, float rho, std::vector<float> &w, const std::vector<std::vector<float>> &predict_points) const final { return this->predict_impl(params, support_vectors, alpha, rho, w, predict_points); }
    [[nodiscard]] std::vector<double> predict_values(const ::plssvm::detail::parameter<double> &params, const std::vector<std::vector<double>> &support_vectors, const std::vector<double> &alpha, double rho, std::vector<double> &w, const std::vector<std::vector<double>> &predict_points) const final { return this->predict_impl(params, support_vectors, alpha, rho, w, predict_points); }

  private:
    void init(const target_platform target) {
        if (target != target_platform::automatic && target != target_platform::gpu_nvidia) {
            throw backend_exception{ fmt::format("Invalid target platform '{}' for the B200 backend!", target) };
        }
        try {
            impl_ = ::plssvm::b200::csvm{ 0 };
        } catch (const ::plssvm::b200::backend_exception &e) {
            throw backend_exception{ e.what() };
        }
        target_ = plssvm::target_platform::gpu_nvidia;
    }

    template <typename T>
    [[nodiscard]] static ::plssvm::b200::parameter<T> convert(const ::plssvm::detail::parameter<T> &p) {
        ::plssvm::b200::parameter<T> out;
        out.kernel_type = static_cast<::plssvm::b200::kernel_function_type>(static_cast<int>(p.kernel_type.value()));
        out.degree = p.degree.value();
        out.gamma = p.gamma.value();
        out.coef0 = p.coef0.value();
        out.cost = p.cost.value();
        return out;
    }
    template <typename T>
    [[nodiscard]] std::pair<std::vector<T>, T> solve_impl(const ::plssvm::detail::parameter<T> &params, const std::vector<std::vector<T>> &A, std::vector<T> b, const T eps, const unsigned long long max_iter) const {
        try {
            return impl_.solve_system_of_linear_equations(convert(params), A, std::move(b), eps, max_iter);
        } catch (const ::plssvm::b200::backend_exception &e) {
            throw backend_exception{ e.what() };
        }
    }
    template <typename T>
    [[nodiscard]] std::vector<T> predict_impl(const ::plssvm::detail::parameter<T> &params, const std::vector<std::vector<T>> &sv, const std::vector<T> &alpha, const T rho, std::vector<T> &w, const std::vector<std::vector<T>> &points) const {
        try {
            return impl_.predict_values(convert(params), sv, alpha, rho, w, points);
        } catch (const ::plssvm::b200::backend_exception &e) {
            throw backend_exception{ e.what() };
        }
    }

    ::plssvm::b200::csvm impl_{ ::plssvm::b200::csvm::deferred{} };
};

}  // namespace plssvm::b200x

namespace plssvm::detail {
/// ~ include/plssvm/backends/CUDA/csvm.hpp:196-197
template <>
struct csvm_backend_exists<b200x::csvm> : std::true_type {};
}  // namespace plssvm::detail

#endif  // PLSSVM_B200_INTEGRATION_CSVM_HPP_
