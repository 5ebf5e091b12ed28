// ORACLE shim (test infrastructure): boost::random::discrete_distribution — min(), max(), probabilities()
// (/root/reference include/cpprob/distributions/utils_discrete.hpp:17-27; the weights are normalised to sum to one, as the
// Boost.Random documentation specifies), param_type, a draw (replay_hook.hpp).
#ifndef CPPROB_REF_SHIM_BOOST_DISCRETE_HPP
#define CPPROB_REF_SHIM_BOOST_DISCRETE_HPP
#include <initializer_list>
#include <istream>
#include <ostream>
#include <random>
#include <vector>
#include <boost/random/replay_hook.hpp>
namespace boost { namespace random {
template<class IntType = int, class WeightType = double>
class discrete_distribution {
public:
    typedef WeightType input_type;
    typedef IntType result_type;
    struct param_type { std::vector<WeightType> p; };
    discrete_distribution() : p_(1, WeightType(1)) {}
    template<class Iter> discrete_distribution(Iter first, Iter last) : p_(first, last) { normalise(); }
    discrete_distribution(std::initializer_list<WeightType> w) : p_(w) { normalise(); }
    result_type min() const { return 0; }
    result_type max() const { return static_cast<result_type>(p_.size() - 1); }
    std::vector<WeightType> probabilities() const { return p_; }
    template<class Engine> result_type operator()(Engine & eng) const
    {
        if (cpprob_ref_shim::replay().active()) return static_cast<result_type>(cpprob_ref_shim::replay().next());
        return static_cast<result_type>(std::discrete_distribution<long>(p_.begin(), p_.end())(eng));
    }
private:
    void normalise()
    {
        WeightType sum = 0;
        for (const WeightType & w : p_) sum += w;
        for (WeightType & w : p_) w /= sum;
    }
    std::vector<WeightType> p_;
};
// streamable like Boost's (the reference's mixture / truncated classes print their members)
template<class I, class W> std::ostream & operator<<(std::ostream & os, const discrete_distribution<I, W> & d) { for (const W & w : d.probabilities()) os << w << ' '; return os; }
template<class I, class W> std::istream & operator>>(std::istream & is, discrete_distribution<I, W> &) { return is; }
}}
#endif
