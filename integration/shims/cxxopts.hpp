// Minimal offline stand-in for cxxopts v3.0.0 (pinned by the reference's CMakeLists.txt:139) — ONLY the API surface the
// reference's command line parsers use (src/plssvm/detail/cmd/parser_{train,predict,scale}.cpp): Options with
// positional_help / show_positional_help / set_width / set_tab_expansion / add_options()(...) / parse_positional / parse /
// help, value<T>() with default_value, ParseResult with count / operator[] / unmatched and OptionValue::as<T> / count.
// Values of non-string types are parsed with operator>> (the reference's enums provide it, e.g. backend_types.cpp:94-112).
#ifndef PLSSVM_B200_CXXOPTS_SHIM_HPP_
#define PLSSVM_B200_CXXOPTS_SHIM_HPP_

#include <any>
#include <cstddef>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace cxxopts {

class OptionException : public std::runtime_error {
  public:
    using std::runtime_error::runtime_error;
};

namespace detail {
template <typename T>
std::any parse_text(const std::string &name, const std::string &text) {
    if constexpr (std::is_same_v<T, std::string>) {
        return std::any{ text };
    } else if constexpr (std::is_same_v<T, bool>) {
        if (text == "true" || text == "1" || text == "True" || text == "TRUE") { return std::any{ true }; }
        if (text == "false" || text == "0" || text == "False" || text == "FALSE") { return std::any{ false }; }
        throw OptionException{ "Argument '" + text + "' failed to parse for option '" + name + "'" };
    } else {
        std::istringstream in{ text };
        T value{};
        in >> value;
        if (in.fail() || (in.peek() != std::char_traits<char>::eof())) { throw OptionException{ "Argument '" + text + "' failed to parse for option '" + name + "'" }; }
        return std::any{ value };
    }
}
}  // namespace detail

class Value : public std::enable_shared_from_this<Value> {
  public:
    virtual ~Value() = default;
    std::shared_ptr<Value> default_value(const std::string &text) {
        has_default_ = true;
        default_ = text;
        return shared_from_this();
    }
    [[nodiscard]] bool has_default() const noexcept { return has_default_; }
    [[nodiscard]] const std::string &get_default() const noexcept { return default_; }
    [[nodiscard]] virtual bool is_boolean() const noexcept = 0;
    [[nodiscard]] virtual std::any parse(const std::string &name, const std::string &text) const = 0;

  private:
    bool has_default_{ false };
    std::string default_{};
};

template <typename T>
class TypedValue final : public Value {
  public:
    [[nodiscard]] bool is_boolean() const noexcept override { return std::is_same_v<T, bool>; }
    [[nodiscard]] std::any parse(const std::string &name, const std::string &text) const override { return detail::parse_text<T>(name, text); }
};

template <typename T>
std::shared_ptr<Value> value() {
    return std::make_shared<TypedValue<T>>();
}

class OptionValue {
  public:
    [[nodiscard]] std::size_t count() const noexcept { return count_; }
    template <typename T>
    [[nodiscard]] const T &as() const {
        if (!value_.has_value()) { throw OptionException{ "Option '" + name_ + "' has no value" }; }
        return std::any_cast<const T &>(value_);
    }

  private:
    friend class Options;
    std::string name_{};
    std::any value_{};
    std::size_t count_{ 0 };
};

class ParseResult {
  public:
    [[nodiscard]] std::size_t count(const std::string &name) const {
        const auto it = values_.find(name);
        return it == values_.end() ? 0 : it->second.count();
    }
    [[nodiscard]] const OptionValue &operator[](const std::string &name) const {
        const auto it = values_.find(name);
        if (it == values_.end()) { throw OptionException{ "Option '" + name + "' does not exist" }; }
        return it->second;
    }
    [[nodiscard]] const std::vector<std::string> &unmatched() const noexcept { return unmatched_; }

  private:
    friend class Options;
    std::map<std::string, OptionValue> values_{};
    std::vector<std::string> unmatched_{};
};

class Options;

class OptionAdder {
  public:
    explicit OptionAdder(Options &options) : options_{ options } {}
    OptionAdder &operator()(const std::string &opts, const std::string &desc, const std::shared_ptr<Value> &val = value<bool>(), const std::string &arg_help = "");

  private:
    Options &options_;
};

class Options {
  public:
    Options(std::string program, std::string help_string = "") : program_{ std::move(program) }, help_string_{ std::move(help_string) } {}

    Options &positional_help(std::string text) {
        positional_help_ = std::move(text);
        return *this;
    }
    Options &show_positional_help() { return *this; }
    Options &set_width(std::size_t) { return *this; }
    Options &set_tab_expansion(bool = true) { return *this; }
    OptionAdder add_options(const std::string & = "") { return OptionAdder{ *this }; }
    void parse_positional(std::initializer_list<std::string> names) { positional_.assign(names.begin(), names.end()); }

    ParseResult parse(const int argc, const char *const *argv) const {
        ParseResult result;
        for (const option &o : options_) {
            OptionValue v;
            v.name_ = o.long_name;
            if (o.value->has_default()) { v.value_ = o.value->parse(o.long_name, o.value->get_default()); }
            result.values_.emplace(o.long_name, std::move(v));
        }
        const auto assign = [&](const option &o, const std::string &text) {
            OptionValue &v = result.values_[o.long_name];
            v.value_ = o.value->parse(o.long_name, text);
            ++v.count_;
        };
        std::size_t next_positional = 0;
        bool only_positional = false;
        for (int i = 1; i < argc; ++i) {
            const std::string arg{ argv[i] };
            if (!only_positional && arg == "--") {
                only_positional = true;
            } else if (!only_positional && arg.size() > 2 && arg[0] == '-' && arg[1] == '-') {
                const std::size_t eq = arg.find('=');
                const std::string name = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
                const option &o = find_long(name);
                if (eq != std::string::npos) {
                    assign(o, arg.substr(eq + 1));
                } else if (o.value->is_boolean()) {
                    assign(o, "true");
                } else {
                    if (i + 1 >= argc) { throw OptionException{ "Option '" + name + "' is missing an argument" }; }
                    assign(o, argv[++i]);
                }
            } else if (!only_positional && arg.size() > 1 && arg[0] == '-' && !(arg[1] >= '0' && arg[1] <= '9') && arg[1] != '.') {
                for (std::size_t c = 1; c < arg.size(); ++c) {
                    const option &o = find_short(arg[c]);
                    if (o.value->is_boolean()) {
                        assign(o, "true");
                    } else {
                        if (c + 1 < arg.size()) {
                            assign(o, arg.substr(c + 1));
                        } else {
                            if (i + 1 >= argc) { throw OptionException{ "Option '" + o.long_name + "' is missing an argument" }; }
                            assign(o, argv[++i]);
                        }
                        break;
                    }
                }
            } else if (next_positional < positional_.size()) {
                assign(find_long(positional_[next_positional++]), arg);
            } else {
                result.unmatched_.push_back(arg);
            }
        }
        return result;
    }
    ParseResult parse(const int argc, char **argv) const { return parse(argc, const_cast<const char *const *>(argv)); }

    [[nodiscard]] std::string help() const {
        std::ostringstream out;
        out << help_string_ << "\nUsage:\n  " << program_ << " [OPTION...] " << positional_help_ << "\n\n";
        for (const option &o : options_) {
            bool is_positional = false;
            for (const std::string &p : positional_) { is_positional = is_positional || p == o.long_name; }
            if (is_positional) { continue; }
            std::string left = "  ";
            left += o.short_name != '\0' ? std::string{ "-" } + o.short_name + ", " : std::string{ "    " };
            left += "--" + o.long_name;
            if (!o.value->is_boolean()) { left += " " + (o.arg_help.empty() ? std::string{ "arg" } : o.arg_help); }
            if (left.size() < 36) { left.resize(36, ' '); }
            out << left << " " << o.desc;
            if (o.value->has_default() && !o.value->is_boolean()) { out << " (default: " << o.value->get_default() << ")"; }
            out << "\n";
        }
        return out.str();
    }

  private:
    friend class OptionAdder;
    struct option {
        char short_name;
        std::string long_name;
        std::string desc;
        std::shared_ptr<Value> value;
        std::string arg_help;
    };
    const option &find_long(const std::string &name) const {
        for (const option &o : options_) {
            if (o.long_name == name) { return o; }
        }
        throw OptionException{ "Option '" + name + "' does not exist" };
    }
    const option &find_short(const char c) const {
        for (const option &o : options_) {
            if (o.short_name == c) { return o; }
        }
        throw OptionException{ std::string{ "Option '" } + c + "' does not exist" };
    }

    std::string program_;
    std::string help_string_;
    std::string positional_help_{};
    std::vector<option> options_{};
    std::vector<std::string> positional_{};
};

inline OptionAdder &OptionAdder::operator()(const std::string &opts, const std::string &desc, const std::shared_ptr<Value> &val, const std::string &arg_help) {
    Options::option o{ '\0', opts, desc, val, arg_help };
    const std::size_t comma = opts.find(',');
    if (comma != std::string::npos) {  // "s,long"
        o.short_name = opts[0];
        o.long_name = opts.substr(comma + 1);
    }
    options_.options_.push_back(std::move(o));
    return *this;
}

}  // namespace cxxopts

#endif  // PLSSVM_B200_CXXOPTS_SHIM_HPP_
