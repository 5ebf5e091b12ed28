// ORACLE shim (test infrastructure): accessor of boost::random::poisson_distribution used by
// /root/reference include/cpprob/distributions/utils_poisson.hpp:17-36.
#ifndef CPPROB_REF_SHIM_BOOST_POISSON_HPP
#define CPPROB_REF_SHIM_BOOST_POISSON_HPP
namespace boost { namespace random {
template<class IntType = int, class RealType = double>
class poisson_distribution {
public:
    typedef IntType result_type;
    typedef RealType input_type;
    explicit poisson_distribution(RealType mean_arg = RealType(1)) : mean_(mean_arg) {}
    RealType mean() const { return mean_; }
private:
    RealType mean_;
};
}}
#endif
