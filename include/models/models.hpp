// cpprob-b200: host symbols of the reference's template models
// (/root/reference include/models/models.hpp:22-35 gaussian_unknown_mean<>, :51-65 Gauss<>,
// :67-80 linear_gaussian_1d<N>, :114-141 hmm<N>) — same names, same signatures.  Each is a stub
// bound to a device functor of models/device_models.hpp (see cpprob/model_binding.hpp): passing it
// to cpprob::inference runs the functor for all particles on the GPU.
#ifndef CPPROB_MODELS_MODELS_HPP
#define CPPROB_MODELS_MODELS_HPP

#include <array>
#include <cstddef>

#include "models/device_models.hpp"

namespace models {

template<class RealType = double>
void gaussian_unknown_mean(const RealType y1, const RealType y2)
{
    const double obs[2] = {static_cast<double>(y1), static_cast<double>(y2)};
    ::cpprob::host_stub<gaussian_unknown_mean_mu_model>(gaussian_unknown_mean_mu_model::name(), obs, 2);
}

template<class RealType = double>
struct Gauss {   // functor form, models.hpp:51-65
    void operator()(const RealType y1, const RealType y2) const { gaussian_unknown_mean<RealType>(y1, y2); }
};

template<std::size_t N>
void linear_gaussian_1d(const std::array<double, N> & observations)
{
    ::cpprob::host_stub<linear_gaussian_1d_model>(linear_gaussian_1d_model::name(), observations.data(), static_cast<int>(N));
}

template<std::size_t N>
void hmm(const std::array<double, N> & observed_states)
{
    ::cpprob::host_stub<hmm_model>(hmm_model::name(), observed_states.data(), static_cast<int>(N));
}
}  // namespace models
#endif  // CPPROB_MODELS_MODELS_HPP
