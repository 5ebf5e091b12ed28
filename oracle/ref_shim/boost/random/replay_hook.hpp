// ORACLE shim (test infrastructure): where the stand-ins for Boost.Random's distributions get their values.
// The reference draws through `distr(get_rng())` (include/cpprob/cpprob.hpp:68-76, global mt19937 seeded from
// random_device, src/cpprob/utils.cpp:16-20); Boost's samplers are not in this image, so the stand-ins either
//   * REPLAY: return the next value of a list the test driver installed (oracle/ref_sis.cpp --replay): the reference's own
//     inference loop then runs on prescribed sampled values, and everything downstream of the draw — log-pdfs, log-weight
//     accumulation, predict routing, address ids, the posterior files — is the reference's own code, or
//   * DRAW with the C++ standard library's distribution of the same family (timing runs; parity is distributional there).
#ifndef CPPROB_REF_SHIM_REPLAY_HOOK_HPP
#define CPPROB_REF_SHIM_REPLAY_HOOK_HPP
#include <cstddef>
#include <stdexcept>
namespace cpprob_ref_shim {
struct replay_state {
    const double * values = nullptr;
    std::size_t n = 0, pos = 0;
    bool active() const { return values != nullptr; }
    double next()
    {
        if (pos >= n) throw std::runtime_error("replay list exhausted");
        return values[pos++];
    }
};
inline replay_state & replay() { static replay_state s; return s; }
}
#endif
