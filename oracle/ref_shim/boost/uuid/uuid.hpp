// ORACLE shim (test infrastructure): boost::uuids::uuid as src/cpprob/socket.cpp:82-85 uses it (a random file name of the
// compile mode's batch dump; never reached on the SIS path)
#ifndef CPPROB_REF_SHIM_BOOST_UUID_HPP
#define CPPROB_REF_SHIM_BOOST_UUID_HPP
#include <cstdint>
#include <ostream>
#include <random>
namespace boost { namespace uuids {
struct uuid { std::uint64_t hi, lo; };
inline std::ostream & operator<<(std::ostream & os, const uuid & u) { return os << std::hex << u.hi << '-' << u.lo << std::dec; }
struct random_generator {
    uuid operator()() { static std::mt19937_64 g{std::random_device{}()}; return uuid{g(), g()}; }
};
}}
#endif
