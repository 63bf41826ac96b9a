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

#include "plssvm/backend_types.hpp"          // plssvm::backend_type
#include "plssvm/csvm.hpp"                  // plssvm::csvm
#include "plssvm/detail/logger.hpp"         // plssvm::detail::log, plssvm::verbosity_level
#include "plssvm/detail/performance_tracker.hpp"  // plssvm::detail::tracking_entry, PLSSVM_DETAIL_PERFORMANCE_TRACKER_ADD_TRACKING_ENTRY
#include "plssvm/detail/type_traits.hpp"    // PLSSVM_REQUIRES
#include "plssvm/exceptions/exceptions.hpp" // plssvm::exception
#include "plssvm/parameter.hpp"             // plssvm::parameter, plssvm::detail::parameter, has_only_parameter_named_args_v
#include "plssvm/target_platforms.hpp"      // plssvm::target_platform

#include "plssvm_b200/csvm.hpp"  // plssvm::b200::csvm (C-ABI adaptor)

#include "fmt/chrono.h"
#include "fmt/core.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <numeric>
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
    // cuda::csvm::init (src/plssvm/backends/CUDA/csvm.cu:48-86): target check, every visible device, the same log lines and tracker entries.
    // PLSSVM_B200_NUM_DEVICES=k restricts the backend to the first k visible devices (CUDA_VISIBLE_DEVICES selects which).
    void init(const target_platform target) {
        if (target != target_platform::automatic && target != target_platform::gpu_nvidia) {
            throw backend_exception{ fmt::format("Invalid target platform '{}' for the B200 backend!", target) };
        }
        ::plssvm::detail::log(verbosity_level::full, "\nUsing B200 as backend.\n");
#if defined(PLSSVM_HAS_B200_BACKEND)
        PLSSVM_DETAIL_PERFORMANCE_TRACKER_ADD_TRACKING_ENTRY((::plssvm::detail::tracking_entry{ "backend", "backend", ::plssvm::backend_type::b200 }));
#endif
        PLSSVM_DETAIL_PERFORMANCE_TRACKER_ADD_TRACKING_ENTRY((::plssvm::detail::tracking_entry{ "backend", "target_platform", ::plssvm::target_platform::gpu_nvidia }));
        int count = 0;
        if (plssvm_b200_device_count(&count) != PLSSVM_B200_OK || count == 0) {
            throw backend_exception{ "B200 backend selected but no CUDA capable devices were found!" };
        }
        if (const char *env = std::getenv("PLSSVM_B200_NUM_DEVICES"); env != nullptr && std::atoi(env) > 0) {
            count = std::min(count, std::atoi(env));
        }
        std::vector<int> devices(static_cast<std::size_t>(count));
        std::iota(devices.begin(), devices.end(), 0);
        try {
            impl_ = ::plssvm::b200::csvm{ devices };
        } catch (const ::plssvm::b200::backend_exception &e) {
            throw backend_exception{ e.what() };
        }
        target_ = plssvm::target_platform::gpu_nvidia;
        ::plssvm::detail::log(verbosity_level::full, "Found {} B200 device(s).\n", ::plssvm::detail::tracking_entry{ "backend", "num_devices", devices.size() });
        ::plssvm::detail::log(verbosity_level::full | verbosity_level::timing, "\n");
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
    // The CG loop runs on the device without a host round trip per iteration, so the per-iteration lines of the reference's loop
    // (gpu_csvm.hpp:569-571, 559-563) are emitted after the solve from the residual history the device recorded; the summary line carries
    // the same tracking entries (gpu_csvm.hpp:637-646).
    template <typename T>
    [[nodiscard]] std::pair<std::vector<T>, T> solve_impl(const ::plssvm::detail::parameter<T> &params, const std::vector<std::vector<T>> &A, std::vector<T> b, const T eps, const unsigned long long max_iter) const {
        try {
            auto result = impl_.solve_system_of_linear_equations(convert(params), A, std::move(b), eps, max_iter);
            const plssvm_b200_timings st = impl_.last_stats();
            const T delta = static_cast<T>(st.cg_residuum), target = static_cast<T>(st.cg_target_residuum);
            const auto avg_time = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::duration<double, std::milli>(st.cg_avg_iteration_ms));
            if (::plssvm::verbosity != verbosity_level::quiet && ((verbosity_level::full | verbosity_level::timing) & ::plssvm::verbosity) != verbosity_level::quiet) {
                const std::vector<double> trace = impl_.last_trace();
                for (std::size_t k = 0; k + 1 < trace.size(); ++k) {
                    ::plssvm::detail::log(verbosity_level::full | verbosity_level::timing, "Start Iteration {} (max: {}) with current residuum {} (target: {}). ", k + 1, max_iter,
                                          static_cast<T>(trace[k]), target);
                    ::plssvm::detail::log(verbosity_level::full | verbosity_level::timing, "Done in {}.\n", avg_time);
                }
            }
            ::plssvm::detail::log(verbosity_level::full | verbosity_level::timing,
                                  "Finished after {}/{} iterations with a residuum of {} (target: {}) and an average iteration time of {}.\n",
                                  ::plssvm::detail::tracking_entry{ "cg", "iterations", static_cast<unsigned long long>(st.cg_iterations) },
                                  ::plssvm::detail::tracking_entry{ "cg", "max_iterations", max_iter },
                                  ::plssvm::detail::tracking_entry{ "cg", "residuum", delta },
                                  ::plssvm::detail::tracking_entry{ "cg", "target_residuum", target },
                                  ::plssvm::detail::tracking_entry{ "cg", "avg_iteration_time", avg_time });
            PLSSVM_DETAIL_PERFORMANCE_TRACKER_ADD_TRACKING_ENTRY((::plssvm::detail::tracking_entry{ "cg", "epsilon", eps }));
            ::plssvm::detail::log(verbosity_level::libsvm, "optimization finished, #iter = {}\n", static_cast<unsigned long long>(st.cg_iterations));
            return result;
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
