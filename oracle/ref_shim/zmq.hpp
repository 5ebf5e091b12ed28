// ORACLE shim (test infrastructure): the names of cppzmq (third party, absent) that /root/reference
// include/cpprob/socket.hpp:12-83 and src/cpprob/socket.cpp mention.  The sockets belong to the compile / CSIS modes;
// the SIS path never opens one.  Any attempt to use them throws.
#ifndef CPPROB_REF_SHIM_ZMQ_HPP
#define CPPROB_REF_SHIM_ZMQ_HPP
#include <cstddef>
#include <stdexcept>
#include <string>
#define ZMQ_REQ 3
#define ZMQ_REP 4
namespace zmq {
class context_t { public: explicit context_t(int = 1) {} };
class message_t {
public:
    message_t() {}
    explicit message_t(size_t) {}
    message_t(const void *, size_t) {}
    void * data() { return nullptr; }
    size_t size() const { return 0; }
};
class socket_t {
public:
    socket_t(context_t &, int) {}
    void bind(const std::string &) { fail(); }
    void bind(const char *) { fail(); }
    void connect(const std::string &) { fail(); }
    void connect(const char *) { fail(); }
    bool send(message_t &, int = 0) { fail(); return false; }
    bool recv(message_t *, int = 0) { fail(); return false; }
private:
    static void fail() { throw std::runtime_error("ZeroMQ is not part of the oracle build (SIS path only)"); }
};
}
#endif
