// cpprob-b200: the reference's example models as per-particle device functors.
//
// Each `*_model` struct restates one model of /root/reference (file:line on each) with the
// statements spelled as calls on the particle context (see include/cpprob/particle.hpp); the same
// source compiles as __device__ code inside the engine's kernels and as host code for the structure
// probe.  The host symbols with the reference's names and signatures — what user code passes to
// cpprob::inference — are stubs bound to these functors; they live in models/models.hpp and
// models/gaussian.hpp (two headers because, as in the reference, both declare a
// `gaussian_unknown_mean`).
//
// `n_scalar_obs` tells the engine how to hand over the flat observation array: k >= 0 means
// f(p, obs[0], ..., obs[k-1]); -1 means f(p, obs_span).  `replayable` marks models whose every
// sampled value is also predicted, in order, so that log_w can be recomputed from a record.
#ifndef CPPROB_MODELS_DEVICE_MODELS_HPP
#define CPPROB_MODELS_DEVICE_MODELS_HPP

#include <array>
#include <cstddef>

#include "cpprob/hd.hpp"
#include "cpprob/distributions/distributions.hpp"
#include "cpprob/model_binding.hpp"
#include "cpprob/particle.hpp"

namespace models {

// -------------------------------------------------------------------------------------------------
// README hello-world — /root/reference src/models/gaussian.cpp:6-17 (mu0 = 1, sigma0 = 1.5,
// sigma = 2, predicts "Mean").  Analytic posterior for x = (3, 4): mean 2.32353, variance 1.05882
// (README.md:118).
// -------------------------------------------------------------------------------------------------
struct gaussian_unknown_mean_model {
    static constexpr int n_scalar_obs = 2;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "gaussian_unknown_mean"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const double x1, const double x2) const
    {
        constexpr double mu0 = 1, sigma0 = 1.5, sigma = 2;      // Hyperparameters

        const ::cpprob::normal_distribution<> prior{mu0, sigma0};
        const double mu = cpprob.sample(prior, true);
        const ::cpprob::normal_distribution<> likelihood{mu, sigma};

        cpprob.observe(likelihood, x1);
        cpprob.observe(likelihood, x2);
        cpprob.predict(mu, "Mean");
    }
};

// -------------------------------------------------------------------------------------------------
// models.hpp variant — /root/reference include/models/models.hpp:22-35 (mu0 = 1, sigma0 = sqrt 5,
// sigma = sqrt 2, predicts "Mu"); this is what `./main --model unk_mean` runs (src/main.cpp:124).
// Posterior N(7.25, 5/6) for y = (8, 9) (thesis p. 85).
// -------------------------------------------------------------------------------------------------
struct gaussian_unknown_mean_mu_model {
    static constexpr int n_scalar_obs = 2;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "gaussian_unknown_mean_mu"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const double y1, const double y2) const
    {
        const ::cpprob::normal_distribution<> prior{1, 2.2360679774997898 /* sqrt(5) */};
        const double mu = cpprob.sample(prior, true);
        const double var = 1.4142135623730951 /* sqrt(2) */;

        const ::cpprob::normal_distribution<> likelihood{mu, var};

        cpprob.observe(likelihood, y1);
        cpprob.observe(likelihood, y2);
        cpprob.predict(mu, "Mu");
    }
};

// -------------------------------------------------------------------------------------------------
// Gaussian linear model — /root/reference include/models/models.hpp:67-80.
// -------------------------------------------------------------------------------------------------
struct linear_gaussian_1d_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "linear_gaussian_1d"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> observations) const
    {
        double state = 0;
        const int n = observations.size();
        int i = 0;
#if !defined(CPPROB_NORMAL_BOX_MULLER)
        // the reference's loop body, four trips at a time: their four standard normals come from one call
        // (philox_stream::next_std_normal_x4 — two Philox blocks and four ziggurat trials side by side)
        for (; i + 4 <= n; i += 4) {
            double z[4];
            cpprob.rng().next_std_normal_x4(z);
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int j = 0; j < 4; ++j) {
                const ::cpprob::normal_of_std<> transition_distr{state, 1, z[j]};
                state = cpprob.sample(transition_distr, true);
                const ::cpprob::normal_distribution<> likelihood{state, 1};
                cpprob.observe(likelihood, observations[i + j]);
                cpprob.predict(state, "State");
            }
        }
#endif
        for (; i < n; ++i) {                       // models.hpp:70-78 as written
            const ::cpprob::normal_distribution<> transition_distr{state, 1};
            state = cpprob.sample(transition_distr, true);
            const ::cpprob::normal_distribution<> likelihood{state, 1};
            cpprob.observe(likelihood, observations[i]);
            cpprob.predict(state, "State");
        }
    }
};

// -------------------------------------------------------------------------------------------------
// Hidden Markov model, 3 states — /root/reference include/models/models.hpp:114-141.
// -------------------------------------------------------------------------------------------------
struct hmm_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr bool draws_normals = false;   // uniform_smallint + discrete only: no ziggurat table needed
    static constexpr const char * name() { return "hmm"; }
    static constexpr int k = 3;
    static constexpr int int_predict_states = k;   // every predicted state lies in [0, 3): the engine needs no pilot for the
                                                   // histogram window and stages four states per byte (staged_kernels.cuh)

    using transition_t = ::cpprob::discrete_distribution<int, double, k>;
    CPPROB_HD static ::cpprob::reg_table<double, k> state_means() { return ::cpprob::reg_table<double, k>{{-1, 0, 1}}; }
    CPPROB_HD static transition_t transition_of(int s)
    {
        const ::cpprob::reg_table<::cpprob::reg_table<double, k>, k> T{{{{0.1, 0.5, 0.4}},
                                                                      {{0.2, 0.2, 0.6}},
                                                                      {{0.15, 0.15, 0.7}}}};
        return transition_t{T[s]};
    }

    // Per-launch table (particle.hpp, "Per-launch model tables").  Everything the reference recomputes for every
    // trace although it does not depend on the particle:
    //   t[2 s], s < 3   : the two sampler thresholds of transition row s, as a pair of 32-bit words
    //   t[8 + 3 i + s]  : logpdf<normal>()(normal(state_mean[s], 1), observed_states[i]) — the same call observe() makes,
    //                     so the same bits (utils_normal_distribution.hpp:38-41 operation order)
    // A step of a trace is then one table-row draw, one 8-byte load and one add.
    static constexpr int kLpdfBase = 8, kLpdfStride = k;
    CPPROB_HD static int scratch_doubles(int n_obs) { return kLpdfBase + kLpdfStride * n_obs; }
    CPPROB_HD static void fill_scratch(double * t, const double * obs, int n_obs, int first, int stride)
    {
        const ::cpprob::reg_table<double, k> state_mean = state_means();
        for (int e = first; e < kLpdfBase + kLpdfStride * n_obs; e += stride) {
            if (e < kLpdfBase) {
                const transition_t tr = transition_of(e < k ? e : 0);
                ::cpprob::detail::pack_u32_pair(t + e, e < k ? tr.thresholds()[0] : 0u, e < k ? tr.thresholds()[1] : 0u);
                continue;
            }
            const int i = (e - kLpdfBase) / kLpdfStride, s = (e - kLpdfBase) % kLpdfStride;
            double v = 0.0;
            if (s < k) {
                const ::cpprob::normal_distribution<> likelihood{state_mean[s], 1};
                v = ::cpprob::logpdf<::cpprob::normal_distribution<>>()(likelihood, obs[i]);
            }
            t[e] = v;
        }
    }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> observed_states) const
    {
        const ::cpprob::uniform_smallint<int> prior{0, 2};
        const double * const tab = cpprob.scratch();
        if (tab != nullptr) {
            // models.hpp:114-141 with the particle-invariant parts read from the table
            const std::uint32_t * const thr = reinterpret_cast<const std::uint32_t *>(tab);
            const double * lp = tab + kLpdfBase;
            auto state = cpprob.sample(prior, true);
            cpprob.predict(state, "State");
            cpprob.increment_log_prob(lp[state]);
            const int n = observed_states.size();
            int i = 1;
            // the six thresholds, read once per trace: rthr[j][s] = threshold j of transition row s
            ::cpprob::reg_table<std::uint32_t, k> rthr[k - 1];
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int s = 0; s < k; ++s) {
                rthr[0].v[s] = thr[2 * s];
                rthr[1].v[s] = thr[2 * s + 1];
            }
            // four steps per Philox block (philox_stream::next_u32x4), then the tail one word at a time
            for (; i + 4 <= n; i += 4) {
                std::uint32_t r[4];
                cpprob.rng().next_u32x4(r);
#if defined(__CUDACC__)
#pragma unroll
#endif
                for (int j = 0; j < 4; ++j) {
                    lp += kLpdfStride;
#if defined(CPPROB_HMM_REG_THRESHOLDS)
                    state = cpprob.sample(::cpprob::reg_discrete_of_word<int, k>(rthr, state, r[j]), true);
#else
                    state = cpprob.sample(::cpprob::table_discrete_of_word<int, k>(thr + 2 * state, r[j]), true);
#endif
                    cpprob.predict(state, "State");
                    cpprob.increment_log_prob(lp[state]);
                }
            }
            for (; i < n; ++i) {
                lp += kLpdfStride;
                state = cpprob.sample(::cpprob::table_discrete_distribution<int, k>(thr + 2 * state), true);
                cpprob.predict(state, "State");
                cpprob.increment_log_prob(lp[state]);
            }
            return;
        }
        const ::cpprob::reg_table<double, k> state_mean = state_means();
        // one transition distribution per row, built once per trace; the reference builds
        // `discrete_distribution{T[state].begin(), T[state].end()}` anew at every step (models.hpp:135), which
        // yields the same three objects
        // (the state is an `int` here, std::size_t in the reference: same values, same records, but 32-bit
        // compares on the device)
        const ::cpprob::reg_table<transition_t, k> transition{{transition_of(0), transition_of(1), transition_of(2)}};
        auto state = cpprob.sample(prior, true);
        cpprob.predict(state, "State");
        auto obs_it = observed_states.begin();
        ::cpprob::normal_distribution<> likelihood{state_mean[state], 1};
        cpprob.observe(likelihood, *obs_it);
        ++obs_it;

        for (; obs_it != observed_states.end(); ++obs_it) {
            state = cpprob.sample(transition[state], true);
            cpprob.predict(state, "State");
            likelihood = ::cpprob::normal_distribution<>{state_mean[state], 1};
            cpprob.observe(likelihood, *obs_it);
        }
    }
};

// -------------------------------------------------------------------------------------------------
// 2-D Gaussian with unknown mean, diagonal covariances — /root/reference include/models/models.hpp:38-49.
// The sampled / predicted value is a vector: recorded as `(0 [m0 m1])` in the .real file.
// -------------------------------------------------------------------------------------------------
struct gaussian_2d_unk_mean_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "gaussian_2d_unk_mean"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> y1) const
    {
        const ::cpprob::multivariate_normal_distribution<double, 2> prior{{1, 2}, {2.2360679774997898 /* sqrt 5 */, 1.7320508075688772 /* sqrt 3 */}};
        const auto mu = cpprob.sample(prior, true);
        const double var = 1.4142135623730951 /* sqrt 2 */;

        const ::cpprob::multivariate_normal_distribution<double, 2> likelihood{mu, var};
        cpprob.observe(likelihood, y1);
        cpprob.predict(mu, "Mu");
    }
};

// -------------------------------------------------------------------------------------------------
// N(1, sqrt 5) prior simulated by rejection sampling from its pdf — models.hpp:82-112.  The number of
// sample statements per trace is data dependent (warp lanes leave the loop at different trips); the
// predict structure is fixed.  `rejection_sampling` is the RAII marker of cpprob.hpp:116-125, which does
// nothing to the weights in SIS (state.cpp:35-46).
// -------------------------------------------------------------------------------------------------
struct rejection_sampling_marker {
    CPPROB_HD rejection_sampling_marker() {}
    CPPROB_HD ~rejection_sampling_marker() {}
};

struct normal_rejection_sampling_model {
    static constexpr int n_scalar_obs = 2;
    static constexpr bool replayable = false;      // the accept/reject draws are not predicted
    static constexpr const char * name() { return "normal_rejection_sampling"; }

    static CPPROB_HD double normal_pdf(double mu, double sigma, double x)
    {
        const double z = (x - mu) / sigma;
        return ::cpprob::dm::exp(-0.5 * z * z) / (sigma * 2.5066282746310002 /* sqrt(2 pi) */);
    }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const double y1, const double y2) const
    {
        const double mu_prior = 1;
        const double sigma_prior = 2.2360679774997898;   // sqrt 5
        const double sigma = 1.4142135623730951;         // sqrt 2

        const double maxval = normal_pdf(mu_prior, sigma_prior, mu_prior);
        const ::cpprob::uniform_real_distribution<> proposal{mu_prior - 20 * sigma_prior, mu_prior + 20 * sigma_prior};
        const ::cpprob::uniform_real_distribution<> accept{0, maxval};
        double mu;

        {rejection_sampling_marker rej{};
            do {
                mu = cpprob.sample(proposal, true);
            } while (cpprob.sample(accept, true) > normal_pdf(mu_prior, sigma_prior, mu));
        }

        const ::cpprob::normal_distribution<> likelihood{mu, sigma};
        cpprob.observe(likelihood, y1);
        cpprob.observe(likelihood, y2);
        cpprob.predict(mu, "Mu");
    }
};

// -------------------------------------------------------------------------------------------------
// Polynomial regression of degree D — /root/reference include/models/poly_adjustment.hpp:17-31,85-95.
// Observations are the points flattened as x0 y0 x1 y1 ...
// -------------------------------------------------------------------------------------------------
template<int D>
struct poly_adjustment_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return D == 1 ? "poly_adjustment_1" : D == 2 ? "poly_adjustment_2" : D == 3 ? "poly_adjustment_3" : "poly_adjustment"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> points) const
    {
        const ::cpprob::normal_distribution<> prior{0, 10};
        double poly[D + 1];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i <= D; ++i) poly[i] = cpprob.sample(prior, true);

        for (int j = 0; j + 1 < points.size(); j += 2) {
            double acc = 0.0;                      // Horner, highest coefficient first (eval_poly :25-29)
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int i = D; i >= 0; --i) acc = acc * points[j] + poly[i];
            const ::cpprob::normal_distribution<> likelihood{acc, 1};
            cpprob.observe(likelihood, points[j + 1]);
        }
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i <= D; ++i) cpprob.predict(poly[i], "Coefficient");
    }
};

// -------------------------------------------------------------------------------------------------
// y = a x + b — poly_adjustment.hpp:60-82 (the Builder parameter only matters for `compile`).
// -------------------------------------------------------------------------------------------------
struct linear_regression_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "linear_regression"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> points) const
    {
        const ::cpprob::normal_distribution<> prior{0, 10};
        const double a = cpprob.sample(prior, true);
        const double b = cpprob.sample(prior, true);

        for (int j = 0; j + 1 < points.size(); j += 2) {
            const ::cpprob::normal_distribution<> likelihood{a * points[j] + b, 1};
            cpprob.observe(likelihood, points[j + 1]);
        }
        cpprob.predict(a, "a");
        cpprob.predict(b, "b");
    }
};

// -------------------------------------------------------------------------------------------------
// One statement of every distribution — /root/reference src/models/models.cpp:13-47.  It uses the
// one-argument predict, whose address the reference takes from the call stack (get_addr(), utils.cpp:71-128 — out of
// scope here).  Known deviation (tests/test_ref_sis_gpu.py): built with -rdynamic, as the reference's CMake does, that
// string carries the call site's code offset, so the reference gives each of the five statements its own id and routes
// the non-const NDArray of the last one to <file>.any; the device has no call-site identity, address() names the
// function once and the vector predict stays in <file>.real.
// -------------------------------------------------------------------------------------------------
struct all_distr_model {
    static constexpr int n_scalar_obs = 2;
    static constexpr bool replayable = false;
    static constexpr const char * name() { return "all_distr"; }
    static constexpr const char * address() { return "[models::all_distr(int, int)]"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, double, double) const
    {
        const ::cpprob::normal_distribution<> normal{1, 2};
        const auto normal_val = cpprob.sample(normal, true);
        cpprob.predict(normal_val);
        cpprob.observe(normal, normal_val);

        const ::cpprob::uniform_smallint<> discrete{2, 7};
        const auto discrete_val = cpprob.sample(discrete, true);
        cpprob.predict(discrete_val);
        cpprob.observe(discrete, discrete_val);

        const ::cpprob::uniform_real_distribution<> rand_unif{2, 9.5};
        const auto rand_unif_val = cpprob.sample(rand_unif, true);
        cpprob.predict(rand_unif_val);
        cpprob.observe(rand_unif, rand_unif_val);

        const ::cpprob::poisson_distribution<> poiss(0.8);
        const auto poiss_val = cpprob.sample(poiss, true);
        cpprob.predict(poiss_val);
        cpprob.observe(poiss, poiss_val);

        const ::cpprob::multivariate_normal_distribution<double, 4> multi{{1, 2, 3, 4}, {2, 1, 5, 3}};
        const auto sample_multi = cpprob.sample(multi, true);
        cpprob.predict(sample_multi);
        cpprob.observe(multi, sample_multi);
    }
};

}  // namespace models
#endif  // CPPROB_MODELS_DEVICE_MODELS_HPP
