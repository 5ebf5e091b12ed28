// cpprob-b200: device distribution layer — POD distribution types, fp64 samplers and log-pdfs.
//
// The reference takes its distribution *types* and *samplers* from Boost.Random (un-vendored,
// pinned to 1.66 by CI) and supplies the log-densities itself as specialisations of
// `cpprob::logpdf<D>` (/root/reference: include/cpprob/distributions/utils_base.hpp:27-28):
//     normal           utils_normal_distribution.hpp:20-45
//     uniform_real     utils_uniform_real.hpp:21-31
//     uniform_smallint utils_uniform_smallint.hpp:17-27
//     discrete         utils_discrete.hpp:17-27
//     poisson          utils_poisson.hpp:17-36
// Boost types are host-only (and Boost is absent from this image), so this header ships trivially
// copyable types with the same accessor names (mean() sigma() a() b() min() max() probabilities()
// result_type) that can live in registers, keeps `cpprob::logpdf<D>` as the customisation point and
// evaluates every log-density in the reference's operation order.  gamma / beta have no reference
// log-pdf; they follow the textbook densities in Boost's parametrisation.
//
// Samplers take a cpprob::philox_stream (see random/philox.hpp for the word-consumption rules).
// The reference pins no sampler output (its RNG is seeded from random_device), so sampler parity is
// distributional; tests/test_distributions_gpu.py checks moments / KS against scipy.
#ifndef CPPROB_DISTRIBUTIONS_HPP
#define CPPROB_DISTRIBUTIONS_HPP

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <stdexcept>

#include "cpprob/hd.hpp"
#include "cpprob/math/dmath.hpp"
#include "cpprob/random/philox.hpp"

namespace cpprob {

// Customisation point, same name and shape as the reference's (utils_base.hpp:27-28).
template<class Distribution> struct logpdf;

// -------------------------------------------------------------------------------------------------
// normal
// -------------------------------------------------------------------------------------------------
template<class RealType = double>
class normal_distribution {
public:
    using result_type = RealType;
    using input_type = RealType;

    CPPROB_HD explicit normal_distribution(RealType mean = 0, RealType sigma = 1) : mean_(mean), sigma_(sigma) {}
    CPPROB_HD RealType mean() const { return mean_; }
    CPPROB_HD RealType sigma() const { return sigma_; }
    CPPROB_HD RealType min() const { return -std::numeric_limits<RealType>::infinity(); }
    CPPROB_HD RealType max() const { return std::numeric_limits<RealType>::infinity(); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        return static_cast<RealType>(rng.next_std_normal() * sigma_ + mean_);
    }

private:
    RealType mean_, sigma_;
};

// A normal whose standard-normal variate was drawn earlier (philox_stream::next_std_normal_x4 hands out four at a
// time): sampling returns z * sigma + mean, the expression normal_distribution uses, without touching the stream.
template<class RealType = double>
class normal_of_std {
public:
    using result_type = RealType;
    using input_type = RealType;

    CPPROB_HD normal_of_std(RealType mean, RealType sigma, RealType z) : mean_(mean), sigma_(sigma), z_(z) {}
    CPPROB_HD RealType mean() const { return mean_; }
    CPPROB_HD RealType sigma() const { return sigma_; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng &) const { return static_cast<RealType>(z_ * sigma_ + mean_); }

private:
    RealType mean_, sigma_, z_;
};

template<class RealType>
struct logpdf<normal_distribution<RealType>> {
    // Operation order of utils_normal_distribution.hpp:26-43 (needed for the 1e-12 replay gate).
    CPPROB_HD RealType operator()(const normal_distribution<RealType> & distr, const RealType & x) const
    {
        const RealType mean = distr.mean();
        const RealType std = distr.sigma();
        if (CPPROB_UNLIKELY(std == 0)) {
            return x == mean ? RealType(0) : -std::numeric_limits<RealType>::infinity();
        }
        if (CPPROB_UNLIKELY(dm::fabs(x) == std::numeric_limits<RealType>::infinity())) {
            return -std::numeric_limits<RealType>::infinity();
        }
        return finite_case(distr, x);
    }
    // The arithmetic of the regular case alone.  Where the exact log-pdf is finite this IS the exact value (same
    // operations, up to FMA contraction); in every special case above it comes out non-finite (x or mean infinite: inf or inf - inf;
    // sigma == 0: x/0 squared plus log 0 = inf - inf), never a wrong finite number.  The fused kernel's fast pass
    // uses it (particle.hpp, `lenient_logpdf`) and recomputes the unit with operator() as soon as a non-finite
    // log-weight shows up, so results are unchanged while the two selects per observe leave the hot loop.
    CPPROB_HD RealType finite_case(const normal_distribution<RealType> & distr, const RealType & x) const
    {
        const RealType std = distr.sigma();
        RealType result = (x - distr.mean()) / std;
        result *= result;
        result += dm::log(2 * dm::pi * std * std);
        result *= -0.5;
        return result;
    }
};

// -------------------------------------------------------------------------------------------------
// uniform_real
// -------------------------------------------------------------------------------------------------
template<class RealType = double>
class uniform_real_distribution {
public:
    using result_type = RealType;
    using input_type = RealType;

    CPPROB_HD explicit uniform_real_distribution(RealType a = 0, RealType b = 1) : a_(a), b_(b) {}
    CPPROB_HD RealType a() const { return a_; }
    CPPROB_HD RealType b() const { return b_; }
    CPPROB_HD RealType min() const { return a_; }
    CPPROB_HD RealType max() const { return b_; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        return static_cast<RealType>(a_ + (b_ - a_) * rng.next_uniform());
    }

private:
    RealType a_, b_;
};

template<class RealType>
struct logpdf<uniform_real_distribution<RealType>> {
    // utils_uniform_real.hpp:24-30
    CPPROB_HD RealType operator()(const uniform_real_distribution<RealType> & distr, const RealType & x) const
    {
        if (x < distr.min() || x > distr.max()) {
            return -std::numeric_limits<RealType>::infinity();
        }
        return -dm::log(distr.b() - distr.a());
    }
};

// -------------------------------------------------------------------------------------------------
// uniform_smallint
// -------------------------------------------------------------------------------------------------
template<class IntType = int>
class uniform_smallint {
public:
    using result_type = IntType;
    using input_type = IntType;

    CPPROB_HD explicit uniform_smallint(IntType min = 0, IntType max = 9) : min_(min), max_(max) {}
    CPPROB_HD IntType a() const { return min_; }
    CPPROB_HD IntType b() const { return max_; }
    CPPROB_HD IntType min() const { return min_; }
    CPPROB_HD IntType max() const { return max_; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        // "small" range: one 32-bit word, multiply-shift (bias <= range * 2^-32, as for Boost's
        // uniform_smallint, which is documented as only approximately uniform).
        const std::uint32_t range = static_cast<std::uint32_t>(max_ - min_) + 1u;
        const std::uint64_t prod = static_cast<std::uint64_t>(rng.next_u32()) * range;
        return static_cast<IntType>(min_ + static_cast<IntType>(prod >> 32));
    }

private:
    IntType min_, max_;
};

template<class IntType>
struct logpdf<uniform_smallint<IntType>> {
    // utils_uniform_smallint.hpp:20-26 (returns double whatever IntType is)
    CPPROB_HD double operator()(const uniform_smallint<IntType> & distr, const IntType & x) const
    {
        if (x < distr.min() || x > distr.max()) {
            return -std::numeric_limits<double>::infinity();
        }
        return -dm::log(distr.max() - distr.min() + 1.0);
    }
};

// -------------------------------------------------------------------------------------------------
// discrete (weights held inline, capacity MaxK; normalised on construction like Boost's)
// -------------------------------------------------------------------------------------------------
// Small constant table held in registers; operator[] is a select chain, so a per-lane index costs
// neither local memory nor a divergent constant-bank load.
template<class T, int N>
struct reg_table {
    T v[N];
    template<class Index>
    CPPROB_HD T operator[](Index i) const
    {
        T r = v[0];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int j = 1; j < N; ++j) if (i == static_cast<Index>(j)) r = v[j];
        return r;
    }
    CPPROB_HD const T * begin() const { return v; }
    CPPROB_HD const T * end() const { return v + N; }
};

template<class WeightType, int MaxK>
struct probability_array {
    WeightType p[MaxK];
    int n;
    CPPROB_HD int size() const { return n; }
    CPPROB_HD WeightType operator[](std::size_t i) const { return p[i]; }
    CPPROB_HD const WeightType * begin() const { return p; }
    CPPROB_HD const WeightType * end() const { return p + n; }
};

template<class IntType = int, class WeightType = double, int MaxK = 8>
class discrete_distribution {
public:
    using result_type = IntType;
    using input_type = WeightType;
    static constexpr int capacity = MaxK;

    CPPROB_HD discrete_distribution() { probs_.n = 1; probs_.p[0] = 1; for (int i = 1; i < MaxK; ++i) probs_.p[i] = 0; set_thresholds(); }

    template<class Iter>
    CPPROB_HD discrete_distribution(Iter first, Iter last) { init(first, last); }

    CPPROB_HD discrete_distribution(std::initializer_list<WeightType> w) { init(w.begin(), w.end()); }

    // from a register table: no pointer into the table is formed, so it never leaves the register file
    template<int N>
    CPPROB_HD discrete_distribution(const reg_table<WeightType, N> & w)
    {
        static_assert(N <= MaxK, "table larger than the distribution's capacity");
        WeightType sum = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) sum += w.v[i];
        probs_.n = N;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < MaxK; ++i) probs_.p[i] = i < N ? w.v[i] : WeightType(0);
        if (sum != WeightType(1)) {          // x / 1 == x exactly: skip the divisions when already normalised
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int i = 0; i < N; ++i) probs_.p[i] = w.v[i] / sum;
        }
        set_thresholds();
    }

    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return static_cast<IntType>(probs_.n - 1); }
    CPPROB_HD const probability_array<WeightType, MaxK> & probabilities() const { return probs_; }

    // Inverse CDF by linear scan on one 32-bit word r: the smallest i with u < p0 + ... + pi, u = (r + 0.5) 2^-32.
    // The comparison is done on integers against thresholds fixed at construction,
    // t_i = ceil((p0 + ... + pi) 2^32 - 0.5)  <=>  (r + 0.5) 2^-32 < p0 + ... + pi  exactly,
    // so drawing costs a few ALU instructions and no FP64 work.
    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        const std::uint32_t r = rng.next_u32();
        int out = probs_.n - 1;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = MaxK - 2; i >= 0; --i) {
            if (i < probs_.n - 1 && r < thr_[i]) out = i;
        }
        return static_cast<IntType>(out);
    }

    // the integer thresholds operator() compares against (thresholds()[i] = t_i, i < size - 1); what a
    // table_discrete_distribution is built from
    CPPROB_HD const std::uint32_t * thresholds() const { return thr_; }

private:
    template<class Iter>
    CPPROB_HD void init(Iter first, Iter last)
    {
        int n = 0;
        WeightType sum = 0;
        for (; first != last && n < MaxK; ++first, ++n) {
            probs_.p[n] = *first;
            sum += probs_.p[n];
        }
        if (first != last) too_many_weights();     // Boost's discrete_distribution has no such limit: never truncate silently
        probs_.n = n;
        for (int i = n; i < MaxK; ++i) probs_.p[i] = WeightType(0);
        if (sum != WeightType(1)) {          // x / 1 == x exactly: skip the divisions when already normalised
            for (int i = 0; i < n; ++i) probs_.p[i] /= sum;
        }
        set_thresholds();
    }

    // more weights than the inline capacity: a device trap / host exception, not a different distribution
    CPPROB_HD static void too_many_weights()
    {
#if CPPROB_ON_DEVICE
        __trap();
#else
        throw std::length_error("cpprob::discrete_distribution: more weights than its capacity MaxK; name a larger MaxK "
                                "(discrete_distribution<IntType, WeightType, MaxK>)");
#endif
    }

    CPPROB_HD void set_thresholds()
    {
        WeightType acc = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < MaxK - 1; ++i) {
            acc += probs_.p[i];
            const WeightType t = acc * WeightType(4294967296.0) - WeightType(0.5);
            // ceil() of a value in [-0.5, 2^32): by truncation plus one when a fraction was dropped
            const WeightType tc = t < WeightType(0) ? WeightType(0) : (t > WeightType(4294967295.0) ? WeightType(4294967295.0) : t);
            const std::uint32_t fl = static_cast<std::uint32_t>(tc);
            thr_[i] = (static_cast<WeightType>(fl) < tc && fl != 0xFFFFFFFFu) ? fl + 1u : fl;
        }
    }

    probability_array<WeightType, MaxK> probs_;
    std::uint32_t thr_[MaxK > 1 ? MaxK - 1 : 1];
};

template<class IntType, class WeightType, int MaxK>
struct logpdf<discrete_distribution<IntType, WeightType, MaxK>> {
    // utils_discrete.hpp:20-26
    CPPROB_HD WeightType operator()(const discrete_distribution<IntType, WeightType, MaxK> & distr, const IntType & x) const
    {
        if (x < distr.min() || x > distr.max()) {
            return -std::numeric_limits<WeightType>::infinity();
        }
        // select instead of a dynamically indexed register array
        WeightType p = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < MaxK; ++i) {
            if (static_cast<IntType>(i) == x) p = distr.probabilities().p[i];
        }
        return dm::log(p);
    }
};

// -------------------------------------------------------------------------------------------------
// table_discrete: the sampler of a discrete_distribution whose K - 1 integer thresholds sit in a table (e.g. one
// row per state of a transition matrix, filled once per launch by Model::fill_scratch).  Draws exactly what the
// discrete_distribution the thresholds came from draws — the smallest i with r < t_i — with K - 1 compares and no
// per-particle set-up.  Sampling only: it carries no probabilities, so it has no logpdf.
// -------------------------------------------------------------------------------------------------
template<class IntType, int K>
class table_discrete_distribution {
public:
    using result_type = IntType;
    using input_type = double;

    CPPROB_HD explicit table_discrete_distribution(const std::uint32_t * thresholds) : thr_(thresholds) {}
    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return static_cast<IntType>(K - 1); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        const std::uint32_t r = rng.next_u32();
        int out = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < K - 1; ++i) out += r >= thr_[i] ? 1 : 0;       // thresholds are non-decreasing
        return static_cast<IntType>(out);
    }

private:
    const std::uint32_t * thr_;
};

// The same draw from a word taken earlier (philox_stream::next_u32x4 hands out four at a time): the stream is not
// touched.
template<class IntType, int K>
class table_discrete_of_word {
public:
    using result_type = IntType;
    using input_type = double;

    CPPROB_HD table_discrete_of_word(const std::uint32_t * thresholds, std::uint32_t word) : thr_(thresholds), r_(word) {}
    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return static_cast<IntType>(K - 1); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng &) const
    {
        int out = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < K - 1; ++i) out += r_ >= thr_[i] ? 1 : 0;
        return static_cast<IntType>(out);
    }

private:
    const std::uint32_t * thr_;
    std::uint32_t r_;
};

// Thresholds of all K rows held in registers: thr[i][state] is threshold i of row `state` (reg_table::operator[] is a
// select chain).  For inner loops that are bound by shared-memory traffic rather than by issue slots: K - 1 selects
// per threshold instead of a table load on the state -> next state dependency chain.
template<class IntType, int K>
class reg_discrete_of_word {
public:
    using result_type = IntType;
    using input_type = double;

    CPPROB_HD reg_discrete_of_word(const reg_table<std::uint32_t, K> (&thr)[K - 1], IntType state, std::uint32_t word) : r_(word)
    {
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < K - 1; ++i) t_[i] = thr[i][state];
    }
    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return static_cast<IntType>(K - 1); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng &) const
    {
        int out = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < K - 1; ++i) out += r_ >= t_[i] ? 1 : 0;
        return static_cast<IntType>(out);
    }

private:
    std::uint32_t t_[K - 1];
    std::uint32_t r_;
};

// -------------------------------------------------------------------------------------------------
// categorical: non-owning view of K probabilities (e.g. one row of a transition matrix).  Not in the
// reference; north_star lists it next to `discrete`.  Probabilities are used as given (no
// normalisation); log-pdf follows the discrete rule.
// -------------------------------------------------------------------------------------------------
template<class IntType = int, class WeightType = double>
class categorical_distribution {
public:
    using result_type = IntType;
    using input_type = WeightType;

    CPPROB_HD categorical_distribution(const WeightType * probs, int k) : probs_(probs), k_(k) {}
    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return static_cast<IntType>(k_ - 1); }
    CPPROB_HD const WeightType * probabilities() const { return probs_; }
    CPPROB_HD int size() const { return k_; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        const WeightType u = (static_cast<WeightType>(rng.next_u32()) + WeightType(0.5)) * WeightType(2.3283064365386962890625e-10);
        WeightType acc = 0;
        int r = k_ - 1;
        for (int i = 0; i < k_ - 1; ++i) {
            acc += probs_[i];
            if (u < acc) { r = i; break; }
        }
        return static_cast<IntType>(r);
    }

private:
    const WeightType * probs_;
    int k_;
};

template<class IntType, class WeightType>
struct logpdf<categorical_distribution<IntType, WeightType>> {
    CPPROB_HD WeightType operator()(const categorical_distribution<IntType, WeightType> & distr, const IntType & x) const
    {
        if (x < distr.min() || x > distr.max()) {
            return -std::numeric_limits<WeightType>::infinity();
        }
        return dm::log(distr.probabilities()[static_cast<std::size_t>(x)]);
    }
};

// -------------------------------------------------------------------------------------------------
// poisson
// -------------------------------------------------------------------------------------------------
template<class IntType = int, class RealType = double>
class poisson_distribution {
public:
    using result_type = IntType;
    using input_type = RealType;

    CPPROB_HD explicit poisson_distribution(RealType mean = 1) : mean_(mean) {}
    CPPROB_HD RealType mean() const { return mean_; }
    CPPROB_HD IntType min() const { return 0; }
    CPPROB_HD IntType max() const { return std::numeric_limits<IntType>::max(); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        if (mean_ < RealType(10)) {
            // inversion by sequential search on the CDF
            const RealType u = rng.next_uniform();
            RealType p = dm::exp(-mean_);
            RealType F = p;
            IntType x = 0;
            while (u > F && x < 1000) {
                ++x;
                p *= mean_ / static_cast<RealType>(x);
                F += p;
            }
            return x;
        }
        // PTRS, transformed rejection with squeeze (W. Hoermann 1993), valid for mean >= 10
        const RealType slam = dm::sqrt(mean_);
        const RealType loglam = dm::log(mean_);
        const RealType b = RealType(0.931) + RealType(2.53) * slam;
        const RealType a = RealType(-0.059) + RealType(0.02483) * b;
        const RealType inv_alpha = RealType(1.1239) + RealType(1.1328) / (b - RealType(3.4));
        const RealType vr = RealType(0.9277) - RealType(3.6224) / (b - RealType(2));
        for (int it = 0; it < 1000; ++it) {
            const RealType U = rng.next_uniform() - RealType(0.5);
            const RealType V = rng.next_uniform();
            const RealType us = RealType(0.5) - dm::fabs(U);
            const RealType kf = dm::floor((RealType(2) * a / us + b) * U + mean_ + RealType(0.43));
            if (us >= RealType(0.07) && V <= vr) return static_cast<IntType>(kf);
            if (kf < 0 || (us < RealType(0.013) && V > us)) continue;
            if (dm::log(V) + dm::log(inv_alpha) - dm::log(a / (us * us) + b)
                <= -mean_ + kf * loglam - dm::lgamma(kf + RealType(1))) {
                return static_cast<IntType>(kf);
            }
        }
        return static_cast<IntType>(mean_);
    }

private:
    RealType mean_;
};

template<class IntType, class RealType>
struct logpdf<poisson_distribution<IntType, RealType>> {
    // utils_poisson.hpp:20-35: explicit sum of logs, not lgamma
    CPPROB_HD RealType operator()(const poisson_distribution<IntType, RealType> & distr, const IntType & x) const
    {
        const RealType l = distr.mean();
        if (l == RealType(0)) {
            return -std::numeric_limits<RealType>::infinity();
        }
        RealType ret = x * dm::log(l) - l;
        for (int i = 1; i <= x; ++i) {
            ret -= dm::log(static_cast<RealType>(i));
        }
        return ret;
    }
};

// -------------------------------------------------------------------------------------------------
// gamma (shape alpha, scale beta — Boost.Random's parametrisation)
// -------------------------------------------------------------------------------------------------
template<class RealType = double>
class gamma_distribution {
public:
    using result_type = RealType;
    using input_type = RealType;

    CPPROB_HD explicit gamma_distribution(RealType alpha = 1, RealType beta = 1) : alpha_(alpha), beta_(beta) {}
    CPPROB_HD RealType alpha() const { return alpha_; }
    CPPROB_HD RealType beta() const { return beta_; }
    CPPROB_HD RealType min() const { return 0; }
    CPPROB_HD RealType max() const { return std::numeric_limits<RealType>::infinity(); }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        // Marsaglia & Tsang (2000); shape < 1 boosted with U^(1/alpha)
        const bool small = alpha_ < RealType(1);
        const RealType a = small ? alpha_ + RealType(1) : alpha_;
        const RealType d = a - RealType(1) / RealType(3);
        const RealType c = RealType(1) / dm::sqrt(RealType(9) * d);
        RealType g = d;
        for (int it = 0; it < 1000; ++it) {
            const RealType x = rng.next_std_normal();
            RealType v = RealType(1) + c * x;
            if (v <= RealType(0)) continue;
            v = v * v * v;
            const RealType u = rng.next_uniform();
            if (dm::log(u) < RealType(0.5) * x * x + d - d * v + d * dm::log(v)) { g = d * v; break; }
        }
        if (small) {
            g *= dm::pow(rng.next_uniform(), RealType(1) / alpha_);
        }
        return g * beta_;
    }

private:
    RealType alpha_, beta_;
};

template<class RealType>
struct logpdf<gamma_distribution<RealType>> {
    CPPROB_HD RealType operator()(const gamma_distribution<RealType> & distr, const RealType & x) const
    {
        const RealType k = distr.alpha();
        const RealType theta = distr.beta();
        if (x < RealType(0)) return -std::numeric_limits<RealType>::infinity();
        if (x == RealType(0)) {
            if (k == RealType(1)) return -dm::log(theta);
            return k < RealType(1) ? std::numeric_limits<RealType>::infinity()
                                   : -std::numeric_limits<RealType>::infinity();
        }
        return (k - RealType(1)) * dm::log(x) - x / theta - dm::lgamma(k) - k * dm::log(theta);
    }
};

// -------------------------------------------------------------------------------------------------
// beta
// -------------------------------------------------------------------------------------------------
template<class RealType = double>
class beta_distribution {
public:
    using result_type = RealType;
    using input_type = RealType;

    CPPROB_HD explicit beta_distribution(RealType alpha = 1, RealType beta = 1) : alpha_(alpha), beta_(beta) {}
    CPPROB_HD RealType alpha() const { return alpha_; }
    CPPROB_HD RealType beta() const { return beta_; }
    CPPROB_HD RealType min() const { return 0; }
    CPPROB_HD RealType max() const { return 1; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        const RealType x = gamma_distribution<RealType>(alpha_, 1)(rng);
        const RealType y = gamma_distribution<RealType>(beta_, 1)(rng);
        return x / (x + y);
    }

private:
    RealType alpha_, beta_;
};

template<class RealType>
struct logpdf<beta_distribution<RealType>> {
    CPPROB_HD RealType operator()(const beta_distribution<RealType> & distr, const RealType & x) const
    {
        const RealType a = distr.alpha();
        const RealType b = distr.beta();
        if (x < RealType(0) || x > RealType(1)) return -std::numeric_limits<RealType>::infinity();
        const RealType log_b = dm::lgamma(a) + dm::lgamma(b) - dm::lgamma(a + b);
        // xlogy-style edges: 0 * log 0 = 0
        const RealType t1 = (a == RealType(1)) ? RealType(0) : (a - RealType(1)) * dm::log(x);
        const RealType t2 = (b == RealType(1)) ? RealType(0) : (b - RealType(1)) * dm::log(RealType(1) - x);
        return t1 + t2 - log_b;
    }
};

// -------------------------------------------------------------------------------------------------
// multivariate normal with diagonal covariance, fixed dimension N
// (/root/reference include/cpprob/distributions/multivariate_normal.hpp:19-311: a vector of
// independent normals; log-pdf utils_multivariate_normal.hpp:20-33 = sum of the component log-pdfs)
// -------------------------------------------------------------------------------------------------
template<class T, int N> struct vecn;   // cpprob/particle.hpp

template<class RealType = double, int N = 2>
class multivariate_normal_distribution {
public:
    using result_type = vecn<RealType, N>;
    using input_type = RealType;

    // The second argument is the COVARIANCE diagonal, as in the reference: every component is stored as
    // normal_distribution(mean, sqrt(covariance)) (multivariate_normal.hpp:167-186, "boost::random::normal uses the
    // standard deviation").  models.hpp:42 passes {sqrt(5), sqrt(3)} there, so its components have sigma 5^(1/4), 3^(1/4).
    // (means..., one covariance for every component) — multivariate_normal.hpp:38-42, :64-68
    template<class RangeMean>
    CPPROB_HD multivariate_normal_distribution(const RangeMean & mean, RealType covariance)
    {
        int i = 0;
        for (auto it = mean.begin(); it != mean.end() && i < N; ++it, ++i) { mean_[i] = *it; sigma_[i] = dm::sqrt(covariance); }
    }
    // (means..., covariances...) — multivariate_normal.hpp:44-61, :70-75
    template<class RangeMean, class RangeCov>
    CPPROB_HD multivariate_normal_distribution(const RangeMean & mean, const RangeCov & covariance)
    {
        int i = 0;
        auto is = covariance.begin();
        for (auto it = mean.begin(); it != mean.end() && i < N; ++it, ++is, ++i) { mean_[i] = *it; sigma_[i] = dm::sqrt(*is); }
    }
    CPPROB_HD multivariate_normal_distribution(std::initializer_list<RealType> mean, std::initializer_list<RealType> covariance)
    {
        int i = 0;
        auto is = covariance.begin();
        for (auto it = mean.begin(); it != mean.end() && i < N; ++it, ++is, ++i) { mean_[i] = *it; sigma_[i] = dm::sqrt(*is); }
    }
    CPPROB_HD multivariate_normal_distribution(std::initializer_list<RealType> mean, RealType covariance)
    {
        int i = 0;
        for (auto it = mean.begin(); it != mean.end() && i < N; ++it, ++i) { mean_[i] = *it; sigma_[i] = dm::sqrt(covariance); }
    }

    CPPROB_HD RealType mean(int i) const { return mean_[i]; }
    CPPROB_HD RealType sigma(int i) const { return sigma_[i]; }
    static constexpr int dimension() { return N; }

    template<class Rng>
    CPPROB_HD result_type operator()(Rng & rng) const
    {
        result_type r;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) r.v[i] = normal_distribution<RealType>(mean_[i], sigma_[i])(rng);
        return r;
    }

private:
    RealType mean_[N], sigma_[N];
};

template<class RealType, int N>
struct logpdf<multivariate_normal_distribution<RealType, N>> {
    template<class Vec>
    CPPROB_HD RealType operator()(const multivariate_normal_distribution<RealType, N> & distr, const Vec & x) const
    {
        RealType ret = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) {
            ret += logpdf<normal_distribution<RealType>>()(normal_distribution<RealType>(distr.mean(i), distr.sigma(i)), x[i]);
        }
        return ret;
    }
};

}  // namespace cpprob
#endif  // CPPROB_DISTRIBUTIONS_HPP
