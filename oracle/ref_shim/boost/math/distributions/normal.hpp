// ORACLE shim (test infrastructure): declarations only.  boost::math::normal_distribution / cdf are named by
// normalise<normal_distribution> (utils_normal_distribution.hpp:93-101), a CSIS-side member the SIS path never instantiates.
#ifndef CPPROB_REF_SHIM_BOOST_MATH_NORMAL_HPP
#define CPPROB_REF_SHIM_BOOST_MATH_NORMAL_HPP
namespace boost { namespace math {
template<class RealType> class normal_distribution;
template<class RealType> RealType cdf(const normal_distribution<RealType> &, const RealType &);
}}
#endif
