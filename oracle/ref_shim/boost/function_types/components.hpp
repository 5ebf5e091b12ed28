// ORACLE shim (test infrastructure): Boost.FunctionTypes as /root/reference include/cpprob/traits.hpp:52-93 and
// call_function.hpp use it — arity, result and parameter types of a function (pointer) or member function pointer (the
// class comes first, as in Boost), the parameter list indexable with boost::mpl::at_c.
#ifndef CPPROB_REF_SHIM_FT_COMPONENTS_HPP
#define CPPROB_REF_SHIM_FT_COMPONENTS_HPP
#include <cstddef>
#include <tuple>
namespace boost { namespace function_types {
namespace detail {
template<class F> struct parts;
template<class R, class... A> struct parts<R(A...)> { typedef R result; typedef std::tuple<A...> params; };
template<class R, class... A> struct parts<R (*)(A...)> : parts<R(A...)> {};
template<class R, class... A> struct parts<R (&)(A...)> : parts<R(A...)> {};
template<class R, class C, class... A> struct parts<R (C::*)(A...)> { typedef R result; typedef std::tuple<C &, A...> params; };
template<class R, class C, class... A> struct parts<R (C::*)(A...) const> { typedef R result; typedef std::tuple<const C &, A...> params; };
template<class F> struct parts<const F> : parts<F> {};
}
template<class F> struct function_arity { static const std::size_t value = std::tuple_size<typename detail::parts<F>::params>::value; };
template<class F> struct result_type { typedef typename detail::parts<F>::result type; };
template<class F> struct parameter_types { typedef typename detail::parts<F>::params as_tuple; };
}}
#endif
