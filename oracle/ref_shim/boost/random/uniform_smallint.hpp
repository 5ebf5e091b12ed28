// ORACLE shim (test infrastructure): accessors of boost::random::uniform_smallint used by
// /root/reference include/cpprob/distributions/utils_uniform_smallint.hpp:17-27,46-49.
#ifndef CPPROB_REF_SHIM_BOOST_SMALLINT_HPP
#define CPPROB_REF_SHIM_BOOST_SMALLINT_HPP
namespace boost { namespace random {
template<class IntType = int>
class uniform_smallint {
public:
    typedef IntType input_type;
    typedef IntType result_type;
    explicit uniform_smallint(IntType min_arg = 0, IntType max_arg = 9) : min_(min_arg), max_(max_arg) {}
    result_type a() const { return min_; }
    result_type b() const { return max_; }
    result_type min() const { return min_; }
    result_type max() const { return max_; }
private:
    IntType min_, max_;
};
}}
#endif
