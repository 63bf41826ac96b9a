// Minimal offline stand-in for fast_float v3.4.0 (pinned by the reference's CMakeLists.txt:166): the one function PLSSVM uses,
// fast_float::from_chars for float / double (include/plssvm/detail/string_conversion.hpp:73-83), forwarded to std::from_chars
// (libstdc++ >= 11 implements the floating-point overloads; both are correctly rounded, so parsed values are identical).
#ifndef PLSSVM_B200_FAST_FLOAT_SHIM_H_
#define PLSSVM_B200_FAST_FLOAT_SHIM_H_

#include <charconv>
#include <system_error>

namespace fast_float {

struct from_chars_result {
    const char *ptr;
    std::errc ec;
};

template <typename T>
from_chars_result from_chars(const char *first, const char *last, T &value) noexcept {
    // fast_float accepts a leading '+', std::from_chars does not
    if (first != last && *first == '+') { ++first; }
    const std::from_chars_result res = std::from_chars(first, last, value);
    return from_chars_result{ res.ptr, res.ec };
}

}  // namespace fast_float

#endif  // PLSSVM_B200_FAST_FLOAT_SHIM_H_
