// cpprob-b200: the public PPL API of CPProb, SIS path, over the B200 engine.
//
// Mirrors /root/reference include/cpprob/cpprob.hpp:
//     inference(StateType, f, observes, n = 50'000, file_name = "posterior", tcp_addr)   :173-203
//     sample(distr, control[, address]) :28-76   observe(distr, x) :79-90   predict(x[, addr]) :92-106
// Same names, namespace, argument order and defaults, so the README program (README.md:102-116)
// compiles unchanged.  What differs underneath:
//   * `inference` does not call `f` n times on the host.  It calls it ONCE; `f` is a host stub that
//     announces its device functor (cpprob/model_binding.hpp), and the n weighted executions, the
//     posterior files and the estimator sums are produced by the GPU engine through the C ABI
//     (include/cpprob_sis.h).  A callable that announces nothing has no device code: `inference`
//     throws std::runtime_error — there is no CPU fallback.
//   * inside a device functor the three statements are calls on the particle context
//     (cpprob/particle.hpp).  The free functions below are their host twins for code that runs
//     outside `inference` (the reference's StateType::dryrun behaviour: sample draws from the prior,
//     observe and predict do nothing).
//   * file_name is a std::string (boost::filesystem::path in the reference; Boost is not a dependency).
//   * the progress line printed every 100 traces (cpprob.hpp:195-197) is dropped: at 10^9 particles it
//     would be 10^7 lines.
// Environment knobs (additions, all optional): CPPROB_SIS_SEED, CPPROB_SIS_DEVICE,
// CPPROB_SIS_EMIT=all|none (none: estimators + .ids + .stats only, no per-particle records),
// CPPROB_SIS_DEVICES=0,1,... (shard the particles over these GPUs of the box, with or without record files).
#ifndef INCLUDE_CPPROB_HPP
#define INCLUDE_CPPROB_HPP

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "cpprob/distributions/distributions.hpp"
#include "cpprob/engine.hpp"
#include "cpprob/model_binding.hpp"
#include "cpprob/state.hpp"

namespace cpprob {

// -------------------------------------------------------------------------------------------------
// call_f_tuple (/root/reference include/cpprob/call_function.hpp:56-65,75-80): invoke a free
// function or a functor on the elements of the observation tuple.
// -------------------------------------------------------------------------------------------------
namespace detail {
template<class F, class Tuple, std::size_t... I>
void call_f_tuple_impl(const F & f, const Tuple & t, std::index_sequence<I...>)
{
    f(std::get<I>(t)...);
}
}  // namespace detail

template<class F, class... Args>
void call_f_tuple(const F & f, const std::tuple<Args...> & args)
{
    detail::call_f_tuple_impl(f, args, std::index_sequence_for<Args...>());
}

// -------------------------------------------------------------------------------------------------
// Host twins of the three statements (dry-run semantics outside `inference`).
// -------------------------------------------------------------------------------------------------
namespace detail {
inline philox_stream & host_stream()
{
    static thread_local philox_keys keys(sis::default_seed());
    static thread_local philox_stream rng(keys, 0);
    return rng;
}
}  // namespace detail

template<class Distribution>
auto sample(Distribution && distr, const bool control = false)
{
    (void)control;
    return distr(detail::host_stream());
}

template<class Distribution, class String>
auto sample(Distribution && distr, const bool control, String &&)
{
    (void)control;
    return distr(detail::host_stream());
}

template<class Distribution>
void observe(Distribution &&, const typename std::decay_t<Distribution>::result_type &) {}

template<class T, class String>
void predict(T &&, String &&) {}

template<class T>
void predict(T &&) {}

// -------------------------------------------------------------------------------------------------
// inference
// -------------------------------------------------------------------------------------------------
namespace detail {
// results of the last inference() on this thread (estimators the engine formed on the device);
// StatsPrinter and user code can read them without re-parsing the files
struct last_run_t {
    bool valid = false;
    cpprob_sis_stats stats;
    std::string file_name;
};
inline last_run_t & last_run()
{
    static thread_local last_run_t r;
    return r;
}
}  // namespace detail

template<class Func, class... Args>
void inference(
        const StateType algorithm,
        const Func & f,
        const std::tuple<Args...> & observes,
        std::size_t n = 50'000,
        const std::string & file_name = "posterior",
        const std::string & tcp_addr = "tcp://127.0.0.1:6666")
{
    static_assert(sizeof...(Args) != 0, "The function has to receive the observed values as parameters.");
    (void)tcp_addr;   // CSIS only

    if (algorithm != StateType::sis) {
        throw std::runtime_error("cpprob-b200 serves StateType::sis only; compile / csis / dryrun belong to the "
                                 "inference-compilation side of CPProb, which this engine leaves untouched");
    }

    // one host call: the stub announces its device model and the flattened observations
    detail::model_announcement who;
    detail::announce_slot() = &who;
    try {
        call_f_tuple(f, observes);
    } catch (...) {
        detail::announce_slot() = nullptr;
        throw;
    }
    detail::announce_slot() = nullptr;
    if (who.name == nullptr) {
        throw std::runtime_error("cpprob::inference: the model passed in is not bound to a device functor "
                                 "(see cpprob/model_binding.hpp); the SIS engine has no CPU fallback");
    }

    int emit = CPPROB_SIS_EMIT_ALL;
    if (const char * e = std::getenv("CPPROB_SIS_EMIT")) {
        if (std::strcmp(e, "none") == 0) emit = CPPROB_SIS_EMIT_NONE;
    }

    detail::last_run_t & last = detail::last_run();
    last.valid = false;
    const std::vector<int> devices = sis::device_list();
    if (devices.size() > 1) {
        // particles sharded over the listed GPUs of this box: estimators only (one NCCL all-gather), or with every rank
        // writing its own particle range of the posterior files
        // (engines, and the NCCL communicator among them, are kept for the next call: sis::cached_engine)
        const std::uint64_t seed = sis::default_seed();
        std::vector<sis::engine *> engines;
        for (int d : devices) {
            engines.push_back(&sis::cached_engine(d));
            engines.back()->set_seed(seed);
        }
        const std::vector<sis::engine *> others(engines.begin() + 1, engines.end());
        const int model = engines[0]->model_id(who.name);
        if (emit == CPPROB_SIS_EMIT_NONE) {
            last.stats = engines[0]->run_multi(others, model, who.obs, n);
            sis::check(cpprob_sis_write_summary(engines[0]->handle(), file_name.c_str(), &last.stats), "cpprob_sis_write_summary");
        } else {
            last.stats = engines[0]->infer_to_files_multi(others, model, who.obs, n, file_name);
        }
    } else {
        sis::engine & engine = sis::cached_engine(devices.size() == 1 ? devices[0] : sis::default_device());
        engine.set_seed(sis::default_seed());
        const int model = engine.model_id(who.name);
        last.stats = engine.infer_to_files(model, who.obs, n, file_name, emit);
    }
    // the stats struct points into engine-owned memory: keep only the scalars
    last.stats.real_mean = last.stats.real_var = last.stats.int_prob = last.stats.sums = nullptr;
    last.stats.int_map = nullptr;
    last.file_name = file_name;
    last.valid = true;
}

}  // end namespace cpprob
#endif  // INCLUDE_CPPROB_HPP
