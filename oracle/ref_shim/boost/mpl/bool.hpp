// ORACLE shim (test infrastructure): boost::mpl::bool_ / true_ / false_
#ifndef CPPROB_REF_SHIM_BOOST_MPL_BOOL_HPP
#define CPPROB_REF_SHIM_BOOST_MPL_BOOL_HPP
namespace boost { namespace mpl {
template<bool C> struct bool_ { static const bool value = C; typedef bool_ type; };
typedef bool_<true> true_;
typedef bool_<false> false_;
}}
#endif
