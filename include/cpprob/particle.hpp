// cpprob-b200: the per-particle execution context — what `sample`, `observe` and `predict` act on.
//
// In the reference the three statements are free function templates that mutate process-global
// statics (/root/reference: include/cpprob/cpprob.hpp:68-76 sample, :79-90 observe, :92-98 predict;
// state behind them in src/cpprob/state.cpp:188-223 and include/cpprob/state.hpp:312-349).  Device
// code has no such statics, so a model is a functor whose first parameter is a `particle<Policy>&`
// and the statements are calls on it:
//
//     struct my_model {
//         template<class P> CPPROB_HD void operator()(P & cpprob, double x1, double x2) const {
//             cpprob::normal_distribution<> prior{1, 1.5};
//             const double mu = cpprob.sample(prior, true);          // was cpprob::sample(prior, true)
//             cpprob.observe(cpprob::normal_distribution<>{mu, 2}, x1);
//             cpprob.predict(mu, "Mean");
//         }
//     };
//
// (free-function spellings cpprob::sample(p, d, control) etc. are provided too).  SIS semantics are
// the reference's: sample = prior draw, adds nothing to log_w (cpprob.hpp:72-74); observe =
// `log_w += logpdf<D>()(d, x)` accumulated in program order from 0.0 (cpprob.hpp:87-89,
// state.cpp:221, trace.hpp:59); predict = record (address id, value), integral values and floating
// values routed to separate record lists (state.hpp:312-326).
//
// The Policy decides where sampled values come from and where predicted values go; the engine
// instantiates each model with the policies in cpprob_b200/csrc/sis_kernels.cuh (register staging,
// SoA rows in HBM, replay from SoA rows) and, on the host, with `probe_policy` below.
#ifndef CPPROB_PARTICLE_HPP
#define CPPROB_PARTICLE_HPP

#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <utility>

#include <string>
#include <unordered_map>
#include <vector>

#include "cpprob/hd.hpp"
#include "cpprob/distributions/distributions.hpp"
#include "cpprob/random/philox.hpp"

namespace cpprob {

// View of the observation array handed to array-style models (`std::array<double, N>` in the
// reference, models.hpp:67-68,114-115; here N is a run-time value).
template<class T>
struct obs_span {
    const T * ptr;
    int n;
    // (base, 32-bit index) instead of a raw pointer: a loop over the observations then carries one 32-bit counter
    // instead of a 64-bit pointer plus a 64-bit end compare (3 instructions less per step of the row kernels)
    struct iterator {
        using value_type = T;
        using reference = const T &;
        using pointer = const T *;
        using difference_type = int;
        const T * base;
        int i;
        CPPROB_HD const T & operator*() const { return base[i]; }
        CPPROB_HD const T * operator->() const { return base + i; }
        CPPROB_HD const T & operator[](int k) const { return base[i + k]; }
        CPPROB_HD iterator & operator++() { ++i; return *this; }
        CPPROB_HD iterator operator++(int) { iterator t = *this; ++i; return t; }
        CPPROB_HD iterator & operator--() { --i; return *this; }
        CPPROB_HD iterator & operator+=(int k) { i += k; return *this; }
        CPPROB_HD iterator operator+(int k) const { return iterator{base, i + k}; }
        CPPROB_HD iterator operator-(int k) const { return iterator{base, i - k}; }
        CPPROB_HD int operator-(const iterator & o) const { return i - o.i; }
        CPPROB_HD bool operator==(const iterator & o) const { return i == o.i; }
        CPPROB_HD bool operator!=(const iterator & o) const { return i != o.i; }
        CPPROB_HD bool operator<(const iterator & o) const { return i < o.i; }
    };
    CPPROB_HD iterator begin() const { return iterator{ptr, 0}; }
    CPPROB_HD iterator end() const { return iterator{ptr, n}; }
    CPPROB_HD const T * data() const { return ptr; }
    CPPROB_HD int size() const { return n; }
    CPPROB_HD const T & operator[](int i) const { return ptr[i]; }
};

// Fixed-size vector value (the reference's NDArray<double> for the vector-valued models,
// /root/reference include/cpprob/ndarray.hpp:26): what a multivariate distribution samples and what
// `predict` records as `[a b ...]`.
template<class T, int N>
struct vecn {
    T v[N];
    static constexpr int size() { return N; }
    CPPROB_HD T & operator[](int i) { return v[i]; }
    CPPROB_HD const T & operator[](int i) const { return v[i]; }
    CPPROB_HD const T * begin() const { return v; }
    CPPROB_HD const T * end() const { return v + N; }
};

namespace detail {
// what logpdf<D> receives: D::result_type for scalar distributions (so that `observe(poisson, 3.0)` converts
// as in the reference), the value itself for vector-valued ones (any indexable range)
template<class D, class Value, bool Scalar = std::is_arithmetic<typename D::result_type>::value>
struct observed { using type = typename D::result_type; };
template<class D, class Value>
struct observed<D, Value, false> { using type = Value; };
}  // namespace detail

namespace detail {
// Policy::lenient_logpdf (absent = false): observe may use logpdf<D>::finite_case where the specialisation has one
template<class Policy, class = void>
struct is_lenient : std::false_type {};
template<class Policy>
struct is_lenient<Policy, decltype(void(Policy::lenient_logpdf))> : std::integral_constant<bool, Policy::lenient_logpdf> {};

// Policy::first_observe_stores (absent = false), see particle::observe
template<class Policy, class = void>
struct first_observe_stores : std::false_type {};
template<class Policy>
struct first_observe_stores<Policy, decltype(void(Policy::first_observe_stores))> : std::integral_constant<bool, Policy::first_observe_stores> {};

template<class D, class X>
CPPROB_HD auto eval_logpdf(std::true_type, const D & d, const X & x, int) -> decltype(logpdf<D>().finite_case(d, x))
{
    return logpdf<D>().finite_case(d, x);
}
template<class D, class X, class Lenient>
CPPROB_HD auto eval_logpdf(Lenient, const D & d, const X & x, long) -> decltype(logpdf<D>()(d, x))
{
    return logpdf<D>()(d, x);
}
}  // namespace detail

template<class Policy>
class particle {
public:
    // `rng` is the stream this particle draws from; the two particles of a stream pair are run one
    // after the other on the same stream object (random/philox.hpp, "particle -> stream map").
    // `scratch`: the model's per-launch table (Model::fill_scratch, see invoke_model below), or nullptr
    CPPROB_HD particle(philox_stream & rng, Policy & policy, const double * scratch = nullptr)
        : rng_(rng), log_w_(0.0), policy_(policy), default_address_("[model]"), scratch_(scratch), observed_(false) {}

    // cpprob::sample(distr, control) — cpprob.hpp:68-76.  `control` is accepted and, as in the
    // reference's SIS branch (:72), has no effect.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class Distribution>
    CPPROB_HD typename Distribution::result_type sample(const Distribution & distr, bool control = false)
    {
        (void)control;
        return policy_.sample(distr, rng_);
    }

    // cpprob::sample(distr, control, address) — cpprob.hpp:28-35; the address is unused in SIS.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class Distribution, class String>
    CPPROB_HD typename Distribution::result_type sample(const Distribution & distr, bool control, const String &)
    {
        (void)control;
        return policy_.sample(distr, rng_);
    }

    // cpprob::observe(distr, x) — cpprob.hpp:79-90 -> StateInfer::increment_log_prob, state.cpp:221.
    template<class Distribution, class Value>
    CPPROB_HD void observe(const Distribution & distr, const Value & x)
    {
        const double lp = detail::eval_logpdf(std::integral_constant<bool, detail::is_lenient<Policy>::value>(), distr,
                                              static_cast<const typename detail::observed<Distribution, Value>::type &>(x), 0);
        // log_w starts at 0.0 (trace.hpp:59).  0.0 + lp == lp, so a policy may ask (`first_observe_stores`) that the
        // first observe stores instead of adding: in a straight-line model `observed_` is resolved at compile time and
        // an FP64 instruction per particle is saved (the fused kernel).  Where observes sit in a loop the flag is a
        // run-time select on every trip, so the default is the reference's plain `+=`.
        increment_log_prob(lp);
    }

    // StateInfer::increment_log_prob (state.cpp:212-223): what `observe` does with the log-density.  Public so that a
    // model with a per-launch table of log-densities (Model::fill_scratch: entries computed by the very same
    // logpdf<D>()(distr, x) calls) can add an entry instead of re-evaluating it for every particle.
    CPPROB_HD void increment_log_prob(const double lp)
    {
        if (detail::first_observe_stores<Policy>::value) {
            log_w_ = observed_ ? log_w_ + lp : lp;
            observed_ = true;
        } else {
            log_w_ += lp;
        }
    }

    // the model's per-launch table, or nullptr when the kernel has none (host probe, tables that do not fit)
    CPPROB_HD const double * scratch() const { return scratch_; }

    // cpprob::predict(x, addr) — cpprob.hpp:92-98 -> StateInfer::add_predict, state.hpp:312-326.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class T, class String,
             typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    CPPROB_HD void predict(T x, const String & addr)
    {
        policy_.predict_int(x, addr);          // with its own width: 32-bit values keep 32-bit bookkeeping on the device
    }

#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class T, class String,
             typename std::enable_if<std::is_floating_point<T>::value, int>::type = 0>
    CPPROB_HD void predict(T x, const String & addr)
    {
        policy_.predict_real(static_cast<double>(x), addr);
    }

    // vector-valued predict: NDArray route of StateInfer::add_predict (state.hpp:328-340), recorded in
    // the .real file as `(id [v0 v1 ...])`
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class T, int N, class String>
    CPPROB_HD void predict(const vecn<T, N> & x, const String & addr)
    {
        policy_.begin_vector(N, addr);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) policy_.predict_real(static_cast<double>(x[i]), addr);
    }

    // cpprob::predict(x) — cpprob.hpp:100-106.  The reference derives the address from the call stack
    // (get_addr, src/cpprob/utils.cpp:71-128), which at function granularity is the same string for every
    // statement of a model: "[<model signature>]".  Device code has no stack trace; the model supplies that
    // string once through set_default_address (invoke_model does it from Model::address()).
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
    template<class T>
    CPPROB_HD void predict(const T & x)
    {
        predict(x, default_address_);
    }
    CPPROB_HD void set_default_address(const char * a) { default_address_ = a; }

    CPPROB_HD double log_w() const { return log_w_; }
    CPPROB_HD philox_stream & rng() { return rng_; }

private:
    philox_stream & rng_;
    double log_w_;
    Policy & policy_;
    const char * default_address_;
    const double * scratch_;
    bool observed_;
};

// Free-function spellings.
template<class P, class D>
CPPROB_HD typename D::result_type sample(particle<P> & p, const D & d, bool control = false) { return p.sample(d, control); }
template<class P, class D, class V>
CPPROB_HD void observe(particle<P> & p, const D & d, const V & x) { p.observe(d, x); }
template<class P, class T, class S>
CPPROB_HD void predict(particle<P> & p, T x, const S & addr) { p.predict(x, addr); }

// -------------------------------------------------------------------------------------------------
// Model invocation: unpack the observation array the way call_f_tuple does for the observation
// tuple (/root/reference: include/cpprob/call_function.hpp:56-65,75-80).
//   Model::n_scalar_obs >= 0 : f(p, obs[0], ..., obs[n-1])      (e.g. gaussian_unknown_mean(x1, x2))
//   Model::n_scalar_obs == -1: f(p, obs_span<double>{obs, n})   (e.g. hmm<N>, linear_gaussian_1d<N>)
// -------------------------------------------------------------------------------------------------
namespace detail {
template<class Model, class P, std::size_t... I>
CPPROB_HD void invoke_scalars(const Model & m, P & p, const double * obs, std::index_sequence<I...>)
{
    m(p, obs[I]...);
}
template<class Model, class P>
CPPROB_HD void invoke_model_impl(const Model & m, P & p, const double * obs, int, std::true_type /*scalars*/)
{
    invoke_scalars(m, p, obs, std::make_index_sequence<static_cast<std::size_t>(Model::n_scalar_obs)>());
}
template<class Model, class P>
CPPROB_HD void invoke_model_impl(const Model & m, P & p, const double * obs, int n, std::false_type)
{
    m(p, obs_span<double>{obs, n});
}
}  // namespace detail

namespace detail {
template<class Model, class P>
CPPROB_HD auto set_address(P & p, int) -> decltype(Model::address(), void()) { p.set_default_address(Model::address()); }
template<class Model, class P>
CPPROB_HD void set_address(P &, long) {}
}  // namespace detail

// -------------------------------------------------------------------------------------------------
// Per-launch model tables.  A model may declare
//     static CPPROB_HD int scratch_doubles(int n_obs);                         // size of its table
//     static CPPROB_HD void fill_scratch(double * t, const double * obs, int n_obs, int first, int stride);
// (entries first, first + stride, ... are filled by the caller: the kernels fill it cooperatively into shared
// memory once per CTA) and read it back through particle::scratch().  Everything particle-invariant that the
// reference recomputes for every trace — log-densities of the observations under each discrete state, sampler
// thresholds — belongs there.  scratch() may be nullptr (host probe, tables beyond the shared-memory budget): the
// model must then compute the same values inline.
// -------------------------------------------------------------------------------------------------
template<class Model, class = void>
struct model_scratch {
    static constexpr bool present = false;
    CPPROB_HD static int doubles(int) { return 0; }
    CPPROB_HD static void fill(double *, const double *, int, int, int) {}
};
template<class Model>
struct model_scratch<Model, decltype(void(Model::scratch_doubles(0)))> {
    static constexpr bool present = true;
    CPPROB_HD static int doubles(int n_obs) { return Model::scratch_doubles(n_obs); }
    CPPROB_HD static void fill(double * t, const double * obs, int n_obs, int first, int stride) { Model::fill_scratch(t, obs, n_obs, first, stride); }
};

template<class Model, class P>
CPPROB_HD void invoke_model(const Model & m, P & p, const double * obs, int n_obs)
{
    detail::set_address<Model>(p, 0);
    detail::invoke_model_impl(m, p, obs, n_obs, std::integral_constant<bool, (Model::n_scalar_obs >= 0)>());
}

// -------------------------------------------------------------------------------------------------
// Host-side structure probe: one execution of the model that records, in program order, the kind
// and address of every predict statement.  Address ids are handed out in first-seen order exactly
// like TraceInfer::register_addr_predict (/root/reference: include/cpprob/trace.hpp:37-41).
// The engine uses the result to lay out the SoA trace rows and to write `<out>.ids`.
// -------------------------------------------------------------------------------------------------
struct predict_slot {
    bool is_int;          // routed to the .int (true) or .real (false) record list
    std::size_t id;       // address id
    std::size_t k;        // occurrence index of this id within the trace (StatsPrinter key)
    std::size_t row;      // (first) row index inside the int / real SoA block
    std::size_t width;    // 1 for a scalar predict, N for a vector predict (N consecutive real rows)
};

struct model_structure {
    std::vector<std::string> ids;          // address strings, index = id
    std::vector<predict_slot> slots;       // program order
    std::size_t n_real = 0, n_int = 0;     // number of real / int predict statements per trace
    std::size_t n_samples = 0;             // number of sample statements per trace
};

class probe_policy {
public:
    explicit probe_policy(model_structure & out) : out_(out) {}

    template<class D>
    typename D::result_type sample(const D & d, philox_stream & rng)
    {
        ++out_.n_samples;
        return d(rng);
    }
    template<class T, class S> void predict_int(T, const S & addr) { add(true, std::string(addr)); }
    template<class S> void predict_real(double, const S & addr)
    {
        if (vector_left_ > 0) {          // component of a vector predict: one more row of the open slot
            ++out_.n_real;
            --vector_left_;
            return;
        }
        add(false, std::string(addr));
    }
    template<class S> void begin_vector(int n, const S & addr)
    {
        add(false, std::string(addr));
        out_.slots.back().width = static_cast<std::size_t>(n);
        --out_.n_real;                   // add() counted one row; the n components count themselves
        vector_left_ = n;
    }

private:
    void add(bool is_int, std::string addr)
    {
        auto it = index_.emplace(addr, out_.ids.size());
        if (it.second) out_.ids.push_back(addr);
        const std::size_t id = it.first->second;
        std::size_t k = 0;
        for (const auto & s : out_.slots) if (s.id == id && s.is_int == is_int) ++k;   // per record file, as StatsPrinter counts
        const std::size_t row = is_int ? out_.n_int++ : out_.n_real++;
        out_.slots.push_back(predict_slot{is_int, id, k, row, 1});
    }

    model_structure & out_;
    std::unordered_map<std::string, std::size_t> index_;
    int vector_left_ = 0;
};

template<class Model>
model_structure probe_model(const Model & m, const double * obs, int n_obs, std::uint64_t seed = 0)
{
    model_structure st;
    probe_policy pol(st);
    const philox_keys keys(seed);
    philox_stream rng(keys, 0);
    particle<probe_policy> p(rng, pol);
    invoke_model(m, p, obs, n_obs);
    return st;
}
}  // namespace cpprob
#endif  // CPPROB_PARTICLE_HPP
