// Stand-in for <fmt/os.h> when building the reference offline against the header-only fmt bundled with torch: that copy
// ships os.h without its compiled part (src/os.cc), so fmt::ostream / fmt::output_file have no definition.  The reference
// only uses `fmt::ostream out = fmt::output_file(name); out.print(fmt, args...)` for writing model / data / scaling files
// (libsvm_model_parsing.hpp:380, libsvm_parsing.hpp:253, arff_parsing.hpp:417, scaling_factors_parsing.hpp:143).
#ifndef PLSSVM_B200_FMT_OS_SHIM_H_
#define PLSSVM_B200_FMT_OS_SHIM_H_

#include <fmt/format.h>

#include <cerrno>
#include <cstdio>
#include <string>
#include <system_error>
#include <utility>

namespace fmt {

class ostream {
  public:
    explicit ostream(const std::string &path) : file_{ std::fopen(path.c_str(), "w") } {
        if (file_ == nullptr) {
            throw std::system_error{ errno, std::generic_category(), "cannot open file " + path };
        }
    }
    ostream(const ostream &) = delete;
    ostream &operator=(const ostream &) = delete;
    ostream(ostream &&other) noexcept : file_{ std::exchange(other.file_, nullptr) } {}
    ~ostream() { close(); }

    template <typename... T>
    void print(format_string<T...> fmt_str, T &&...args) {
        const std::string text = fmt::format(fmt_str, std::forward<T>(args)...);
        std::fwrite(text.data(), 1, text.size(), file_);
    }
    void flush() {
        if (file_ != nullptr) { std::fflush(file_); }
    }
    void close() {
        if (file_ != nullptr) {
            std::fclose(file_);
            file_ = nullptr;
        }
    }

  private:
    std::FILE *file_;
};

inline ostream output_file(const std::string &path) { return ostream{ path }; }

}  // namespace fmt

#endif  // PLSSVM_B200_FMT_OS_SHIM_H_
