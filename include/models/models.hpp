// cpprob-b200: host symbols of the reference's template models
// (/root/reference include/models/models.hpp:22-35 gaussian_unknown_mean<>, :51-65 Gauss<>,
// :67-80 linear_gaussian_1d<N>, :114-141 hmm<N>) — same names, same signatures.  Each is a stub
// bound to a device functor of models/device_models.hpp (see cpprob/model_binding.hpp): passing it
// to cpprob::inference runs the functor for all particles on the GPU.
#ifndef CPPROB_MODELS_MODELS_HPP
#define CPPROB_MODELS_MODELS_HPP

#include <array>
#include <cstddef>
#include <utility>
#include <vector>

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

// models.hpp:38-49
template<class RealType = double>
void gaussian_2d_unk_mean(const std::vector<RealType> y1)
{
    const std::vector<double> obs(y1.begin(), y1.end());
    ::cpprob::host_stub<gaussian_2d_unk_mean_model>(gaussian_2d_unk_mean_model::name(), obs.data(), static_cast<int>(obs.size()));
}

// models.hpp:82-112
template<class RealType = double>
void normal_rejection_sampling(const RealType y1, const RealType y2)
{
    const double obs[2] = {static_cast<double>(y1), static_cast<double>(y2)};
    ::cpprob::host_stub<normal_rejection_sampling_model>(normal_rejection_sampling_model::name(), obs, 2);
}

// poly_adjustment.hpp:85-95 (D <= 3 has device code)
template<std::size_t D, std::size_t N, class RealType = double>
void poly_adjustment(const std::array<std::array<RealType, 2>, N> & points)
{
    static_assert(D >= 1 && D <= 3, "device functors are instantiated for polynomial degrees 1..3");
    double obs[2 * N];
    for (std::size_t i = 0; i < N; ++i) { obs[2 * i] = points[i][0]; obs[2 * i + 1] = points[i][1]; }
    ::cpprob::host_stub<poly_adjustment_model<static_cast<int>(D)>>(poly_adjustment_model<static_cast<int>(D)>::name(), obs, static_cast<int>(2 * N));
}

// poly_adjustment.hpp:60-82 without the Builder argument (it only feeds `compile`)
template<class RealType = double>
void linear_regression(const std::vector<std::pair<RealType, RealType>> & points)
{
    std::vector<double> obs;
    for (const auto & pt : points) { obs.push_back(pt.first); obs.push_back(pt.second); }
    ::cpprob::host_stub<linear_regression_model>(linear_regression_model::name(), obs.data(), static_cast<int>(obs.size()));
}

// src/models/models.cpp:13-47
inline void all_distr(int a, int b)
{
    const double obs[2] = {static_cast<double>(a), static_cast<double>(b)};
    ::cpprob::host_stub<all_distr_model>(all_distr_model::name(), obs, 2);
}
}  // namespace models
#endif  // CPPROB_MODELS_MODELS_HPP
