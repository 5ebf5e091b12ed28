// ORACLE shim (test infrastructure): boost::random::normal_distribution as /root/reference uses it — mean(), sigma()
// (include/cpprob/distributions/utils_normal_distribution.hpp:20-45), param_type, and a draw (see replay_hook.hpp).
#ifndef CPPROB_REF_SHIM_BOOST_NORMAL_HPP
#define CPPROB_REF_SHIM_BOOST_NORMAL_HPP
#include <istream>
#include <ostream>
#include <limits>
#include <random>
#include <boost/random/replay_hook.hpp>
namespace boost { namespace random {
template<class RealType = double>
class normal_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    struct param_type { RealType mean, sigma; };
    explicit normal_distribution(RealType mean_arg = RealType(0), RealType sigma_arg = RealType(1)) : mean_(mean_arg), sigma_(sigma_arg) {}
    RealType mean() const { return mean_; }
    RealType sigma() const { return sigma_; }
    RealType min() const { return -std::numeric_limits<RealType>::infinity(); }
    RealType max() const { return std::numeric_limits<RealType>::infinity(); }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        if (cpprob_ref_shim::replay().active()) return static_cast<result_type>(cpprob_ref_shim::replay().next());
        return std::normal_distribution<RealType>(mean_, sigma_)(eng);
    }
private:
    RealType mean_, sigma_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class R> std::ostream & operator<<(std::ostream & os, const normal_distribution<R> & d) { return os << d.mean() << ' ' << d.sigma(); }
template<class R> std::istream & operator>>(std::istream & is, normal_distribution<R> & d) { R m, s; if (is >> m >> s) d = normal_distribution<R>(m, s); return is; }
}
using random::normal_distribution;        // src/models/gaussian.cpp spells it boost::normal_distribution<>
}
#endif
