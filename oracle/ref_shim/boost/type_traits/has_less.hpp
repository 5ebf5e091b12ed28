// ORACLE shim (test infrastructure): stands in for Boost's has_less, the one Boost header
// /root/reference include/cpprob/utils.hpp:13 needs.  Detects `a < b` by expression SFINAE.
#ifndef CPPROB_REF_SHIM_HAS_LESS_HPP
#define CPPROB_REF_SHIM_HAS_LESS_HPP
#include <type_traits>
#include <utility>
namespace boost {
namespace shim_detail {
template<class T> auto has_less_impl(int) -> decltype(void(std::declval<const T &>() < std::declval<const T &>()), std::true_type{});
template<class T> std::false_type has_less_impl(...);
}
template<class T> struct has_less : decltype(shim_detail::has_less_impl<T>(0)) {};
}
#endif
