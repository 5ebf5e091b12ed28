// ORACLE shim (test infrastructure): boost::throw_exception(e) throws e
#ifndef CPPROB_REF_SHIM_BOOST_THROW_EXCEPTION_HPP
#define CPPROB_REF_SHIM_BOOST_THROW_EXCEPTION_HPP
namespace boost { template<class E> [[noreturn]] inline void throw_exception(const E & e) { throw e; } }
#endif
