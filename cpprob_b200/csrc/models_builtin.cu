// cpprob-b200: kernel instantiations + registry entries for the reference's example models
// (include/models/device_models.hpp).
#include "model_vtable.cuh"
#include "models/device_models.hpp"

CPPROB_SIS_REGISTER_MODEL(models::gaussian_unknown_mean_model)
CPPROB_SIS_REGISTER_MODEL(models::gaussian_unknown_mean_mu_model)
CPPROB_SIS_REGISTER_MODEL(models::linear_gaussian_1d_model)
CPPROB_SIS_REGISTER_MODEL(models::hmm_model)
