// A user model outside the built-in set, compiled into its own shared object and registered with the engine at load
// time: theta ~ Beta(2, 2), every observed flip ~ Bernoulli(theta), predict theta.
// The conjugate posterior Beta(2 + heads, 2 + tails) is what tests/test_plugin_gpu.py checks against.
#include "model_vtable.cuh"

struct coin_model {
    static constexpr int n_scalar_obs = -1;          // f(p, obs_span)
    static constexpr bool replayable = true;
    static constexpr const char * name() { return "coin"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> flips) const
    {
        const double theta = cpprob.sample(::cpprob::beta_distribution<>{2, 2}, true);
        const double w[2] = {1 - theta, theta};
        const ::cpprob::discrete_distribution<int, double, 2> flip{w, w + 2};
        for (const double f : flips) cpprob.observe(flip, static_cast<int>(f));
        cpprob.predict(theta, "Theta");
    }
};

CPPROB_SIS_REGISTER_MODEL(coin_model)
