// ORACLE shim (test infrastructure): declaration only.  /root/reference include/cpprob/traits.hpp:83 names
// boost::mpl::at_c inside a template that the post-processing path never instantiates.
#ifndef CPPROB_REF_SHIM_MPL_AT_HPP
#define CPPROB_REF_SHIM_MPL_AT_HPP
namespace boost { namespace mpl { template<class Seq, long N> struct at_c; } }
#endif
