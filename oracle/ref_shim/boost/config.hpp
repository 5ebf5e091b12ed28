// ORACLE shim (test infrastructure): the Boost.Config macros /root/reference include/cpprob/any.hpp names.
#ifndef CPPROB_REF_SHIM_BOOST_CONFIG_HPP
#define CPPROB_REF_SHIM_BOOST_CONFIG_HPP
#define BOOST_WORKAROUND(symbol, test) 0
#define BOOST_DEDUCED_TYPENAME typename
#define BOOST_HAS_RVALUE_REFS
#endif
