// ORACLE shim (test infrastructure): boost::lexical_cast<std::string>(x) through a stream (src/cpprob/socket.cpp:85)
#ifndef CPPROB_REF_SHIM_BOOST_LEXICAL_CAST_HPP
#define CPPROB_REF_SHIM_BOOST_LEXICAL_CAST_HPP
#include <sstream>
#include <string>
namespace boost { template<class To, class From> To lexical_cast(const From & x) { std::ostringstream s; s << x; return To(s.str()); } }
#endif
