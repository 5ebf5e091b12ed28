// ORACLE shim (test infrastructure): boost::detail::sp_typeinfo = std::type_info
#ifndef CPPROB_REF_SHIM_BOOST_SP_TYPEINFO_HPP
#define CPPROB_REF_SHIM_BOOST_SP_TYPEINFO_HPP
#include <typeinfo>
namespace boost { namespace detail { typedef std::type_info sp_typeinfo; } }
#define BOOST_SP_TYPEID(T) typeid(T)
#endif
