// cpprob-b200: kernel instantiations + registry entries for the reference's example models
// (include/models/device_models.hpp).
#include "model_vtable.cuh"
#include "models/device_models.hpp"

CPPROB_SIS_REGISTER_MODEL(models::gaussian_unknown_mean_model)
CPPROB_SIS_REGISTER_MODEL(models::gaussian_unknown_mean_mu_model)
CPPROB_SIS_REGISTER_MODEL(models::linear_gaussian_1d_model)
CPPROB_SIS_REGISTER_MODEL(models::hmm_model)
CPPROB_SIS_REGISTER_MODEL(models::gaussian_2d_unk_mean_model)
CPPROB_SIS_REGISTER_MODEL(models::normal_rejection_sampling_model)
CPPROB_SIS_REGISTER_MODEL(models::poly_adjustment_model<1>)
CPPROB_SIS_REGISTER_MODEL(models::poly_adjustment_model<2>)
CPPROB_SIS_REGISTER_MODEL(models::poly_adjustment_model<3>)
CPPROB_SIS_REGISTER_MODEL(models::linear_regression_model)
CPPROB_SIS_REGISTER_MODEL(models::all_distr_model)
