// ORACLE shim (test infrastructure): the part of boost::random::discrete_distribution that
// /root/reference include/cpprob/distributions/utils_discrete.hpp:17-27 touches — min(), max(), probabilities().
// Semantics per the Boost.Random documentation: the weights are normalised to sum to one; probabilities() returns them.
#ifndef CPPROB_REF_SHIM_BOOST_DISCRETE_HPP
#define CPPROB_REF_SHIM_BOOST_DISCRETE_HPP
#include <vector>
namespace boost { namespace random {
template<class IntType = int, class WeightType = double>
class discrete_distribution {
public:
    typedef WeightType input_type;
    typedef IntType result_type;
    template<class Iter> discrete_distribution(Iter first, Iter last) : p_(first, last)
    {
        WeightType sum = 0;
        for (const WeightType & w : p_) sum += w;
        for (WeightType & w : p_) w /= sum;
    }
    result_type min() const { return 0; }
    result_type max() const { return static_cast<result_type>(p_.size() - 1); }
    std::vector<WeightType> probabilities() const { return p_; }
private:
    std::vector<WeightType> p_;
};
}}
#endif
