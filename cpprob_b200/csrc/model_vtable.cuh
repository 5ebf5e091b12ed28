// cpprob-b200: per-model launch table.  The engine core (sis_capi.cu) is model-agnostic; every
// model functor is turned into a `cpprob_sis_model_vtable` of host launchers for its kernel
// instantiations.  Built-in models are registered in models_builtin.cu; a user model is registered
// from its own .cu with
//
//     #include "model_vtable.cuh"
//     CPPROB_SIS_REGISTER_MODEL(my_namespace::my_model)
//
// compiled with nvcc (-gencode arch=compute_100a,code=sm_100a) into a shared object that links
// against libcpprob_sis.so.
#ifndef CPPROB_B200_MODEL_VTABLE_CUH
#define CPPROB_B200_MODEL_VTABLE_CUH

#include <cuda_runtime.h>

#include "cpprob_sis.h"
#include "sis_kernels.cuh"

extern "C" {
struct cpprob_sis_model_vtable {
    int abi_version;
    const char * name;
    int n_scalar_obs;           // >= 0: model takes that many scalar observations; -1: an array
    int replayable;
    // host-side structure probe; `out` is a cpprob::model_structure*
    int (*probe)(const double * obs, int n_obs, unsigned long long seed, void * out);
    cudaError_t (*launch_pilot)(cudaStream_t s, const cpprob::philox_keys * keys, const double * obs, int n_obs, int n_pilot, double * out);
    // nr = register-staged real predict slots (1, 2 or 4)
    cudaError_t (*launch_fused)(cudaStream_t s, int grid, int nr, const cpprob::engine::run_args * a);
    cudaError_t (*launch_rows)(cudaStream_t s, int grid, const cpprob::engine::run_args * a);
    cudaError_t (*launch_replay)(cudaStream_t s, int grid, const double * obs, int n_obs, const double * real_rows,
                                 const int * int_rows, unsigned long long stride, unsigned long long n, double * logw_out);
    // resident CTAs per SM: which = 0 fused(nr=1), 1 fused(nr=2), 2 fused(nr=4), 3 rows
    int (*occupancy)(int which);
};
}

namespace cpprob {
namespace engine {

// Dynamic shared memory of a model's kernels: the ziggurat table, if the model draws normals.  More than the 48 KB a
// kernel may use without asking, hence the attribute (set at every launch: it is per device, and cheap).
template<class Model>
constexpr unsigned model_smem() { return model_draws_normals<Model>::value ? zig::kSharedBytes : 0u; }

template<class Kernel>
inline cudaError_t allow_smem(Kernel k, unsigned bytes)
{
    return bytes > 48u * 1024u ? cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) : cudaSuccess;
}

template<class Model>
struct model_launchers {
    static int probe(const double * obs, int n_obs, unsigned long long seed, void * out)
    {
        *static_cast<model_structure *>(out) = probe_model(Model{}, obs, n_obs, seed);
        return 0;
    }
    static cudaError_t pilot(cudaStream_t s, const philox_keys * keys, const double * obs, int n_obs, int n_pilot, double * out)
    {
        if (cudaError_t err = allow_smem(k_pilot<Model>, model_smem<Model>())) return err;
        k_pilot<Model><<<(n_pilot + 511) / 512, kBlock, model_smem<Model>(), s>>>(*keys, obs, n_obs, n_pilot, out);
        return cudaGetLastError();
    }
    static cudaError_t fused(cudaStream_t s, int grid, int nr, const run_args * a)
    {
        constexpr unsigned smem = model_smem<Model>();
        if (nr <= 1) {
            if (cudaError_t err = allow_smem(k_sis_fused<Model, 1>, smem)) return err;
            k_sis_fused<Model, 1><<<grid, fused_block(1), smem, s>>>(*a);
        } else if (nr == 2) {
            if (cudaError_t err = allow_smem(k_sis_fused<Model, 2>, smem)) return err;
            k_sis_fused<Model, 2><<<grid, fused_block(2), smem, s>>>(*a);
        } else {
            if (cudaError_t err = allow_smem(k_sis_fused<Model, 4>, smem)) return err;
            k_sis_fused<Model, 4><<<grid, fused_block(4), smem, s>>>(*a);
        }
        return cudaGetLastError();
    }
    static cudaError_t rows(cudaStream_t s, int grid, const run_args * a)
    {
        if (cudaError_t err = allow_smem(k_sis_rows<Model>, model_smem<Model>())) return err;
        k_sis_rows<Model><<<grid, kBlock, model_smem<Model>(), s>>>(*a);
        return cudaGetLastError();
    }
    static cudaError_t replay(cudaStream_t s, int grid, const double * obs, int n_obs, const double * real_rows,
                              const int * int_rows, unsigned long long stride, unsigned long long n, double * logw_out)
    {
        k_replay<Model><<<grid, kBlock, 0, s>>>(obs, n_obs, real_rows, int_rows, stride, n, logw_out);
        return cudaGetLastError();
    }
    static int occupancy(int which)
    {
        int n = 0;
        cudaError_t err;
        constexpr unsigned smem = model_smem<Model>();
        switch (which) {
        case 0:
            allow_smem(k_sis_fused<Model, 1>, smem);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sis_fused<Model, 1>, fused_block(1), smem);
            break;
        case 1:
            allow_smem(k_sis_fused<Model, 2>, smem);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sis_fused<Model, 2>, fused_block(2), smem);
            break;
        case 2:
            allow_smem(k_sis_fused<Model, 4>, smem);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sis_fused<Model, 4>, fused_block(4), smem);
            break;
        default:
            allow_smem(k_sis_rows<Model>, smem);
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sis_rows<Model>, kBlock, smem);
            break;
        }
        return err == cudaSuccess ? n : 0;
    }
    static const cpprob_sis_model_vtable * vtable()
    {
        static const cpprob_sis_model_vtable vt = {
            CPPROB_SIS_ABI_VERSION, Model::name(), Model::n_scalar_obs, Model::replayable ? 1 : 0,
            &probe, &pilot, &fused, &rows, &replay, &occupancy};
        return &vt;
    }
};

template<class Model>
struct model_registrar {
    model_registrar() { cpprob_sis_register_model(model_launchers<Model>::vtable()); }
};

}  // namespace engine
}  // namespace cpprob

#define CPPROB_SIS_CAT2(a, b) a##b
#define CPPROB_SIS_CAT(a, b) CPPROB_SIS_CAT2(a, b)
#define CPPROB_SIS_REGISTER_MODEL(...) \
    static ::cpprob::engine::model_registrar<__VA_ARGS__> CPPROB_SIS_CAT(cpprob_sis_registrar_, __LINE__);

#endif  // CPPROB_B200_MODEL_VTABLE_CUH
