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
        for (const auto obs : observations) {
            const ::cpprob::normal_distribution<> transition_distr{state, 1};
            state = cpprob.sample(transition_distr, true);
            const ::cpprob::normal_distribution<> likelihood{state, 1};
            cpprob.observe(likelihood, obs);
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
    static constexpr const char * name() { return "hmm"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> observed_states) const
    {
        constexpr int k = 3;
        const ::cpprob::reg_table<double, k> state_mean{{-1, 0, 1}};
        const ::cpprob::reg_table<::cpprob::reg_table<double, k>, k> T{{{{0.1, 0.5, 0.4}},
                                                                      {{0.2, 0.2, 0.6}},
                                                                      {{0.15, 0.15, 0.7}}}};
        const ::cpprob::uniform_smallint<std::size_t> prior{0, 2};
        auto state = cpprob.sample(prior, true);
        cpprob.predict(state, "State");
        auto obs_it = observed_states.begin();
        ::cpprob::normal_distribution<> likelihood{state_mean[state], 1};
        cpprob.observe(likelihood, *obs_it);
        ++obs_it;

        for (; obs_it != observed_states.end(); ++obs_it) {
            const ::cpprob::discrete_distribution<std::size_t, double, k> transition_distr{T[state]};   // {T[state].begin(), T[state].end()}
            state = cpprob.sample(transition_distr, true);
            cpprob.predict(state, "State");
            likelihood = ::cpprob::normal_distribution<>{state_mean[state], 1};
            cpprob.observe(likelihood, *obs_it);
        }
    }
};

}  // namespace models
#endif  // CPPROB_MODELS_DEVICE_MODELS_HPP
