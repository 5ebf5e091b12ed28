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

// A second plugin model, for the limits that must fail loudly: it rolls a four-sided die per observation and predicts the
// face.  `dice<K>` DECLARES that its integral predicts lie in [0, K) (int_predict_states): with K = 4 that is true — the
// engine then needs no pilot for the histogram window and stages four faces per byte — with K = 3 it is a lie, and every
// run must end with CPPROB_SIS_ERANGE instead of a histogram that silently drops the fourth face.
template<int Declared>
struct dice_model {
    static constexpr int n_scalar_obs = -1;
    static constexpr bool replayable = true;
    static constexpr bool draws_normals = false;
    static constexpr int int_predict_states = Declared;
    static constexpr const char * name() { return Declared == 4 ? "dice4" : "dice_lying"; }

    template<class P>
    CPPROB_HD void operator()(P & cpprob, const ::cpprob::obs_span<double> rolls) const
    {
        const ::cpprob::uniform_smallint<int> die{0, 3};
        for (const double r : rolls) {
            const int face = cpprob.sample(die, true);
            cpprob.observe(::cpprob::normal_distribution<>{static_cast<double>(face), 1}, r);
            cpprob.predict(face, "Face");
        }
    }
};

CPPROB_SIS_REGISTER_MODEL(dice_model<4>)
CPPROB_SIS_REGISTER_MODEL(dice_model<3>)
