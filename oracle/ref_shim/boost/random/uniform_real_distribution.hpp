// ORACLE shim (test infrastructure): accessors of boost::random::uniform_real_distribution used by
// /root/reference include/cpprob/distributions/utils_uniform_real.hpp:21-31 (a(), b(), min(), max()).
#ifndef CPPROB_REF_SHIM_BOOST_UNIFORM_REAL_HPP
#define CPPROB_REF_SHIM_BOOST_UNIFORM_REAL_HPP
namespace boost { namespace random {
template<class RealType = double>
class uniform_real_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    explicit uniform_real_distribution(RealType min_arg = RealType(0), RealType max_arg = RealType(1)) : min_(min_arg), max_(max_arg) {}
    RealType a() const { return min_; }
    RealType b() const { return max_; }
    RealType min() const { return min_; }
    RealType max() const { return max_; }
private:
    RealType min_, max_;
};
}}
#endif
