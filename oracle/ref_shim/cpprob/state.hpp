// ORACLE shim (test infrastructure).  /root/reference include/cpprob/postprocess/stats_printer.hpp:16 includes
// cpprob/state.hpp but uses nothing from it; the real header drags in FlatBuffers and ZeroMQ, which are absent from
// this image.  This empty stand-in is found first on the include path, so stats_printer.hpp itself compiles unmodified.
#ifndef CPPROB_REF_SHIM_STATE_HPP
#define CPPROB_REF_SHIM_STATE_HPP
#endif
