// ORACLE shim (test infrastructure): boost::mpl::at_c over the parameter list of function_types/components.hpp
// (/root/reference include/cpprob/traits.hpp:83)
#ifndef CPPROB_REF_SHIM_MPL_AT_HPP
#define CPPROB_REF_SHIM_MPL_AT_HPP
#include <tuple>
namespace boost { namespace mpl {
template<class Seq, long N> struct at_c { typedef typename std::tuple_element<N, typename Seq::as_tuple>::type type; };
}}
#endif
