// Force-included (-include) before every reference translation unit when building the reference offline against fmt >= 10
// (this image ships header-only fmt 12 inside torch/include; the reference pins fmt 8.1.1, CMakeLists.txt:228).
// fmt >= 9 no longer formats types through operator<< implicitly and fmt >= 11 moved fmt::join to <fmt/ranges.h>; the
// reference relies on both (e.g. parameter.hpp:225, libsvm_model_parsing.hpp:164).  This prelude restores that behaviour
// for the reference's own streamable types.  It contains no PLSSVM code.
#ifndef PLSSVM_B200_FMT_PRELUDE_HPP_
#define PLSSVM_B200_FMT_PRELUDE_HPP_

#include <fmt/chrono.h>
#include <fmt/color.h>
#include <fmt/core.h>
#include <fmt/format.h>
#include <fmt/ostream.h>
#include <fmt/ranges.h>

#include <ostream>
#include <type_traits>
#include <utility>

namespace plssvm {
template <typename T>
class default_value;
namespace detail {
template <typename T>
struct parameter;
class execution_range;
template <typename T>
struct tracking_entry;
namespace cmd {
class parser_train;
class parser_predict;
class parser_scale;
}  // namespace cmd
}  // namespace detail
}  // namespace plssvm

// fmt::localtime was removed in fmt 12 (used by src/plssvm/detail/utility.cpp:20 for a time stamp string only)
#include <ctime>
namespace fmt {
inline std::tm localtime(const std::time_t time) {
    std::tm tm{};
    ::localtime_r(&time, &tm);
    return tm;
}
}  // namespace fmt

namespace plssvm_b200_shim {
template <typename T, typename = void>
struct is_streamable : std::false_type {};
template <typename T>
struct is_streamable<T, std::void_t<decltype(std::declval<std::ostream &>() << std::declval<const T &>())>> : std::true_type {};
}  // namespace plssvm_b200_shim

// scoped enums with a user-provided operator<< (backend_type, kernel_function_type, target_platform, file_format_type, ...)
template <typename T>
struct fmt::formatter<T, char, std::enable_if_t<std::is_enum_v<T> && !std::is_convertible_v<T, int> && plssvm_b200_shim::is_streamable<T>::value>> : fmt::ostream_formatter {};
template <typename T>
struct fmt::formatter<plssvm::default_value<T>> : fmt::ostream_formatter {};
template <typename T>
struct fmt::formatter<plssvm::detail::parameter<T>> : fmt::ostream_formatter {};
template <>
struct fmt::formatter<plssvm::detail::execution_range> : fmt::ostream_formatter {};
template <typename T>
struct fmt::formatter<plssvm::detail::tracking_entry<T>> : fmt::ostream_formatter {};

template <>
struct fmt::formatter<plssvm::detail::cmd::parser_train> : fmt::ostream_formatter {};
template <>
struct fmt::formatter<plssvm::detail::cmd::parser_predict> : fmt::ostream_formatter {};
template <>
struct fmt::formatter<plssvm::detail::cmd::parser_scale> : fmt::ostream_formatter {};

#endif  // PLSSVM_B200_FMT_PRELUDE_HPP_
