// ORACLE shim (test infrastructure): boost::random::uniform_real_distribution — a(), b(), min(), max()
// (/root/reference include/cpprob/distributions/utils_uniform_real.hpp:21-31), param_type, a draw (replay_hook.hpp).
#ifndef CPPROB_REF_SHIM_BOOST_UNIFORM_REAL_HPP
#define CPPROB_REF_SHIM_BOOST_UNIFORM_REAL_HPP
#include <istream>
#include <ostream>
#include <random>
#include <boost/random/replay_hook.hpp>
namespace boost { namespace random {
template<class RealType = double>
class uniform_real_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    struct param_type { RealType a, b; };
    explicit uniform_real_distribution(RealType min_arg = RealType(0), RealType max_arg = RealType(1)) : min_(min_arg), max_(max_arg) {}
    RealType a() const { return min_; }
    RealType b() const { return max_; }
    RealType min() const { return min_; }
    RealType max() const { return max_; }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        if (cpprob_ref_shim::replay().active()) return static_cast<result_type>(cpprob_ref_shim::replay().next());
        return std::uniform_real_distribution<RealType>(min_, max_)(eng);
    }
private:
    RealType min_, max_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class R> std::ostream & operator<<(std::ostream & os, const uniform_real_distribution<R> & d) { return os << d.a() << ' ' << d.b(); }
template<class R> std::istream & operator>>(std::istream & is, uniform_real_distribution<R> & d) { R a, b; if (is >> a >> b) d = uniform_real_distribution<R>(a, b); return is; }
}}
#endif
