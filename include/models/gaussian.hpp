// cpprob-b200: host symbol of the README hello-world model
// (/root/reference: include/models/gaussian.hpp:1-3, src/models/gaussian.cpp:6-17).
// The body lives in models::gaussian_unknown_mean_model (models/device_models.hpp); this symbol is
// the stub that user code hands to cpprob::inference (README.md:102-116).
#ifndef CPPROB_MODELS_GAUSSIAN_HPP
#define CPPROB_MODELS_GAUSSIAN_HPP

#include "models/device_models.hpp"

namespace models {
inline void gaussian_unknown_mean(const double x1, const double x2)
{
    const double obs[2] = {x1, x2};
    ::cpprob::host_stub<gaussian_unknown_mean_model>(gaussian_unknown_mean_model::name(), obs, 2);
}
}  // namespace models
#endif  // CPPROB_MODELS_GAUSSIAN_HPP
