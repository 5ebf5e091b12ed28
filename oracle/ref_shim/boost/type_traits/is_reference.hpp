// ORACLE shim (test infrastructure): boost::is_reference = std::is_reference
#ifndef CPPROB_REF_SHIM_BOOST_IS_REFERENCE_HPP
#define CPPROB_REF_SHIM_BOOST_IS_REFERENCE_HPP
#include <type_traits>
namespace boost { template<class T> struct is_reference : std::is_reference<T> {}; }
#endif
