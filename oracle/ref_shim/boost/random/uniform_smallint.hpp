// ORACLE shim (test infrastructure): boost::random::uniform_smallint — a(), b(), min(), max()
// (/root/reference include/cpprob/distributions/utils_uniform_smallint.hpp:17-27,46-49), param_type, a draw (replay_hook.hpp).
#ifndef CPPROB_REF_SHIM_BOOST_SMALLINT_HPP
#define CPPROB_REF_SHIM_BOOST_SMALLINT_HPP
#include <istream>
#include <ostream>
#include <random>
#include <boost/random/replay_hook.hpp>
namespace boost { namespace random {
template<class IntType = int>
class uniform_smallint {
public:
    typedef IntType input_type;
    typedef IntType result_type;
    struct param_type { IntType a, b; };
    explicit uniform_smallint(IntType min_arg = 0, IntType max_arg = 9) : min_(min_arg), max_(max_arg) {}
    result_type a() const { return min_; }
    result_type b() const { return max_; }
    result_type min() const { return min_; }
    result_type max() const { return max_; }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        if (cpprob_ref_shim::replay().active()) return static_cast<result_type>(cpprob_ref_shim::replay().next());
        return std::uniform_int_distribution<IntType>(min_, max_)(eng);
    }
private:
    IntType min_, max_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class I> std::ostream & operator<<(std::ostream & os, const uniform_smallint<I> & d) { return os << d.a() << ' ' << d.b(); }
template<class I> std::istream & operator>>(std::istream & is, uniform_smallint<I> & d) { I a, b; if (is >> a >> b) d = uniform_smallint<I>(a, b); return is; }
}}
#endif
