// ORACLE shim (test infrastructure): boost::remove_reference = std::remove_reference
#ifndef CPPROB_REF_SHIM_BOOST_REMOVE_REFERENCE_HPP
#define CPPROB_REF_SHIM_BOOST_REMOVE_REFERENCE_HPP
#include <type_traits>
namespace boost { template<class T> struct remove_reference : std::remove_reference<T> {}; }
#endif
