// ORACLE shim (test infrastructure): declaration only.  /root/reference include/cpprob/traits.hpp:52-93 names
// boost::function_types::result_type inside templates that the post-processing path never instantiates.
#ifndef CPPROB_REF_SHIM_FT_result_type_HPP
#define CPPROB_REF_SHIM_FT_result_type_HPP
namespace boost { namespace function_types { template<class F> struct result_type; } }
#endif
