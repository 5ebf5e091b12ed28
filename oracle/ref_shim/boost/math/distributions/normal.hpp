// ORACLE shim (test infrastructure): boost::math::normal_distribution with pdf and cdf, as /root/reference uses them
// (include/models/models.hpp:82-112 normal_rejection_sampling: pdf; utils_normal_distribution.hpp:93-101: cdf of a CSIS member).
#ifndef CPPROB_REF_SHIM_BOOST_MATH_NORMAL_HPP
#define CPPROB_REF_SHIM_BOOST_MATH_NORMAL_HPP
#include <cmath>
#include <boost/math/constants/constants.hpp>
namespace boost { namespace math {
template<class RealType = double>
class normal_distribution {
public:
    typedef RealType value_type;
    explicit normal_distribution(RealType mean = 0, RealType sd = 1) : mean_(mean), sd_(sd) {}
    RealType mean() const { return mean_; }
    RealType standard_deviation() const { return sd_; }
private:
    RealType mean_, sd_;
};
typedef normal_distribution<double> normal;
template<class RealType> RealType pdf(const normal_distribution<RealType> & d, const RealType & x)
{
    const RealType z = (x - d.mean()) / d.standard_deviation();
    return std::exp(-z * z / 2) / (d.standard_deviation() * std::sqrt(2 * constants::pi<RealType>()));
}
template<class RealType> RealType cdf(const normal_distribution<RealType> & d, const RealType & x)
{
    return std::erfc(-(x - d.mean()) / (d.standard_deviation() * std::sqrt(RealType(2)))) / 2;
}
}}
#endif
