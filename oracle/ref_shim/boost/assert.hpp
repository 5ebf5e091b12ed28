// ORACLE shim (test infrastructure)
#ifndef CPPROB_REF_SHIM_BOOST_ASSERT_HPP
#define CPPROB_REF_SHIM_BOOST_ASSERT_HPP
#include <cassert>
#define BOOST_ASSERT(x) assert(x)
#endif
