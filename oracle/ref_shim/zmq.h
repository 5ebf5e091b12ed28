// ORACLE shim (test infrastructure): the two ZeroMQ constants src/cpprob/socket.cpp names; see zmq.hpp
#include <zmq.hpp>
