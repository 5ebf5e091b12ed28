// ORACLE shim (test infrastructure): boost::math::constants::pi<T>() — for double, the correctly rounded value of pi, which is
// what Boost.Math returns (used by /root/reference include/cpprob/distributions/utils_normal_distribution.hpp:40).
#ifndef CPPROB_REF_SHIM_BOOST_MATH_CONSTANTS_HPP
#define CPPROB_REF_SHIM_BOOST_MATH_CONSTANTS_HPP
namespace boost { namespace math { namespace constants {
template<class T> inline constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
}}}
#endif
