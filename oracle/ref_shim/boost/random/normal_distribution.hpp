// ORACLE shim (test infrastructure): accessors of boost::random::normal_distribution used by
// /root/reference include/cpprob/distributions/utils_normal_distribution.hpp:20-45 (mean(), sigma()).
#ifndef CPPROB_REF_SHIM_BOOST_NORMAL_HPP
#define CPPROB_REF_SHIM_BOOST_NORMAL_HPP
namespace boost { namespace random {
template<class RealType = double>
class normal_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    explicit normal_distribution(RealType mean_arg = RealType(0), RealType sigma_arg = RealType(1)) : mean_(mean_arg), sigma_(sigma_arg) {}
    RealType mean() const { return mean_; }
    RealType sigma() const { return sigma_; }
private:
    RealType mean_, sigma_;
};
}}
#endif
