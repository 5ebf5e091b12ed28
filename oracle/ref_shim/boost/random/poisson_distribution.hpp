// ORACLE shim (test infrastructure): boost::random::poisson_distribution — mean()
// (/root/reference include/cpprob/distributions/utils_poisson.hpp:17-36), param_type, a draw (replay_hook.hpp).
#ifndef CPPROB_REF_SHIM_BOOST_POISSON_HPP
#define CPPROB_REF_SHIM_BOOST_POISSON_HPP
#include <istream>
#include <ostream>
#include <limits>
#include <random>
#include <boost/random/replay_hook.hpp>
namespace boost { namespace random {
template<class IntType = int, class RealType = double>
class poisson_distribution {
public:
    typedef IntType result_type;
    typedef RealType input_type;
    struct param_type { RealType mean; };
    explicit poisson_distribution(RealType mean_arg = RealType(1)) : mean_(mean_arg) {}
    RealType mean() const { return mean_; }
    IntType min() const { return 0; }
    IntType max() const { return std::numeric_limits<IntType>::max(); }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        if (cpprob_ref_shim::replay().active()) return static_cast<result_type>(cpprob_ref_shim::replay().next());
        return static_cast<result_type>(std::poisson_distribution<long>(mean_)(eng));
    }
private:
    RealType mean_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class I, class R> std::ostream & operator<<(std::ostream & os, const poisson_distribution<I, R> & d) { return os << d.mean(); }
template<class I, class R> std::istream & operator>>(std::istream & is, poisson_distribution<I, R> & d) { R m; if (is >> m) d = poisson_distribution<I, R>(m); return is; }
}}
#endif
