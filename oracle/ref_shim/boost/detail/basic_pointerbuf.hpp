// ORACLE shim (test infrastructure): included by /root/reference src/cpprob/socket.cpp:13, nothing of it is used.
