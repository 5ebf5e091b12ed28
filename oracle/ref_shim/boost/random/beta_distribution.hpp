// ORACLE shim (test infrastructure): boost::random::beta_distribution, named by /root/reference
// include/cpprob/distributions/min_max_continuous.hpp (a CSIS proposal family): accessors and a draw by two gammas.
#ifndef CPPROB_REF_SHIM_BOOST_BETA_HPP
#define CPPROB_REF_SHIM_BOOST_BETA_HPP
#include <istream>
#include <ostream>
#include <random>
namespace boost { namespace random {
template<class RealType = double>
class beta_distribution {
public:
    typedef RealType input_type;
    typedef RealType result_type;
    struct param_type { RealType alpha, beta; };
    explicit beta_distribution(RealType a = RealType(1), RealType b = RealType(1)) : a_(a), b_(b) {}
    RealType alpha() const { return a_; }
    RealType beta() const { return b_; }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        const RealType x = std::gamma_distribution<RealType>(a_, 1)(eng), y = std::gamma_distribution<RealType>(b_, 1)(eng);
        return x / (x + y);
    }
private:
    RealType a_, b_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class R> std::ostream & operator<<(std::ostream & os, const beta_distribution<R> & d) { return os << d.alpha() << ' ' << d.beta(); }
template<class R> std::istream & operator>>(std::istream & is, beta_distribution<R> & d) { R a, b; if (is >> a >> b) d = beta_distribution<R>(a, b); return is; }
}}
#endif
