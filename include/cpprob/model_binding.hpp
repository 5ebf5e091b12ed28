// cpprob-b200: how a host model symbol (what the user passes to cpprob::inference) is tied to the
// device functor that the engine actually runs.
//
// In the reference `inference` simply calls the model function n times on the host
// (/root/reference: include/cpprob/cpprob.hpp:194-201, call_f_tuple in call_function.hpp:56-80).
// Here the host symbol is a stub: when `inference` invokes it ONCE, with the user's observation
// tuple, the stub *announces* the name of its device functor and the flattened observations, and
// the engine runs that functor for all particles on the GPU.  Called directly (outside
// `inference`) a stub behaves like the reference under StateType::dryrun: one prior execution on
// the host with nothing recorded.
#ifndef CPPROB_MODEL_BINDING_HPP
#define CPPROB_MODEL_BINDING_HPP

#include <cstdint>
#include <vector>

#include "cpprob/particle.hpp"

namespace cpprob {
namespace detail {

struct model_announcement {
    const char * name = nullptr;
    std::vector<double> obs;
};

inline model_announcement *& announce_slot()
{
    static thread_local model_announcement * slot = nullptr;
    return slot;
}

struct dryrun_policy {
    template<class D> typename D::result_type sample(const D & d, philox_stream & rng) { return d(rng); }
    template<class S> void predict_int(long long, const S &) {}
    template<class S> void predict_real(double, const S &) {}
    template<class S> void begin_vector(int, const S &) {}
};

inline std::uint64_t & dryrun_counter()
{
    static thread_local std::uint64_t n = 0;
    return n;
}

}  // namespace detail

template<class Model>
void host_stub(const char * device_model_name, const double * obs, int n_obs)
{
    if (detail::model_announcement * a = detail::announce_slot()) {
        a->name = device_model_name;
        a->obs.assign(obs, obs + n_obs);
        return;
    }
    detail::dryrun_policy pol;
    const philox_keys keys(static_cast<std::uint64_t>(0));
    philox_stream rng(keys, detail::dryrun_counter()++);
    particle<detail::dryrun_policy> p(rng, pol);
    invoke_model(Model{}, p, obs, n_obs);
}

}  // namespace cpprob
#endif  // CPPROB_MODEL_BINDING_HPP
