// Minimal offline stand-in for the `igor` named-argument library (https://github.com/bluescarni/igor, commit a5224c6 pinned by
// the reference's CMakeLists.txt:198) — ONLY the API surface PLSSVM v2.0.0 uses: IGOR_MAKE_NAMED_ARGUMENT, `name = value`,
// igor::parser{...} with has / operator() / has_unnamed_arguments / has_duplicates / has_other_than, and the free
// igor::has_unnamed_arguments<Args...>() / igor::has_other_than<Args...>(names...).  Written from the usage in
// include/plssvm/parameter.hpp:35-72,217-265, csvm.hpp:262-294, csvm_factory.hpp:93-101; contains no arithmetic.
#ifndef PLSSVM_B200_IGOR_SHIM_HPP_
#define PLSSVM_B200_IGOR_SHIM_HPP_

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace igor {

namespace detail {

template <typename Tag, typename T>
struct tagged_ref {
    using tag_type = Tag;
    T &&value;
};

template <typename T>
struct is_tagged : std::false_type {};
template <typename Tag, typename T>
struct is_tagged<tagged_ref<Tag, T>> : std::true_type {};
template <typename T>
inline constexpr bool is_tagged_v = is_tagged<std::remove_cv_t<std::remove_reference_t<T>>>::value;

template <typename T, typename = void>
struct tag_of {
    using type = void;
};
template <typename T>
struct tag_of<T, std::enable_if_t<is_tagged_v<T>>> {
    using type = typename std::remove_cv_t<std::remove_reference_t<T>>::tag_type;
};
template <typename T>
using tag_of_t = typename tag_of<T>::type;

}  // namespace detail

template <typename Tag>
struct named_argument {
    using tag_type = Tag;
    template <typename T>
    constexpr detail::tagged_ref<Tag, T> operator=(T &&value) const {
        return detail::tagged_ref<Tag, T>{ std::forward<T>(value) };
    }
};

#define IGOR_MAKE_NAMED_ARGUMENT(name) inline constexpr auto name = ::igor::named_argument<struct name##_igor_tag> {}

template <typename... Args>
constexpr bool has_unnamed_arguments() {
    return (... || !detail::is_tagged_v<Args>);
}

namespace detail {
template <typename Tag, typename... Names>
inline constexpr bool tag_in_v = (... || std::is_same_v<Tag, typename Names::tag_type>);
}  // namespace detail

template <typename... Args, typename... Names>
constexpr bool has_other_than(const Names &...) {
    // true if some tagged argument carries a tag that is not among Names
    return (... || (detail::is_tagged_v<Args> && !detail::tag_in_v<detail::tag_of_t<Args>, Names...>) );
}

template <typename... Args>
class parser {
  public:
    constexpr explicit parser(Args &&...args) : args_{ std::forward<Args>(args)... } {}

    template <typename Name>
    constexpr bool has(const Name &) const {
        return (... || std::is_same_v<detail::tag_of_t<Args>, typename Name::tag_type>);
    }
    constexpr bool has_unnamed_arguments() const { return ::igor::has_unnamed_arguments<Args...>(); }
    template <typename... Names>
    constexpr bool has_other_than(const Names &...names) const {
        return ::igor::has_other_than<Args...>(names...);
    }
    constexpr bool has_duplicates() const { return duplicates_impl<Args...>(); }

    template <typename Name>
    constexpr decltype(auto) operator()(const Name &) const {
        return get_impl<typename Name::tag_type, 0>();
    }

  private:
    template <typename First = void, typename... Rest>
    static constexpr bool duplicates_impl() {
        if constexpr (sizeof...(Rest) == 0) {
            return false;
        } else {
            return (detail::is_tagged_v<First> && (... || std::is_same_v<detail::tag_of_t<First>, detail::tag_of_t<Rest>>) ) || duplicates_impl<Rest...>();
        }
    }
    template <typename Tag, std::size_t I>
    constexpr decltype(auto) get_impl() const {
        static_assert(I < sizeof...(Args), "igor shim: named argument not present");
        using arg_t = std::tuple_element_t<I, std::tuple<Args...>>;
        if constexpr (std::is_same_v<detail::tag_of_t<arg_t>, Tag>) {
            return (std::get<I>(args_).value);
        } else {
            return get_impl<Tag, I + 1>();
        }
    }

    std::tuple<Args &&...> args_;
};

template <typename... Args>
parser(Args &&...) -> parser<Args...>;

}  // namespace igor

#endif  // PLSSVM_B200_IGOR_SHIM_HPP_
