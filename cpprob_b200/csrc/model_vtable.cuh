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
#include "staged_kernels.cuh"

extern "C" {
struct cpprob_sis_model_vtable {
    int abi_version;
    const char * name;
    int n_scalar_obs;           // >= 0: model takes that many scalar observations; -1: an array
    int replayable;
    // host-side structure probe; `out` is a cpprob::model_structure*
    int (*probe)(const double * obs, int n_obs, unsigned long long seed, void * out);
    // fin: null = per-tile maxima only; else the last CTA folds them into out[0] = m_ref (sis_kernels.cuh, pilot_finalize)
    cudaError_t (*launch_pilot)(cudaStream_t s, const cpprob::philox_keys * keys, const double * obs, int n_obs, int n_pilot, double * out,
                                const cpprob::engine::pilot_finalize * fin);
    // nr = register-staged real predict slots (1, 2 or 4)
    cudaError_t (*launch_fused)(cudaStream_t s, int grid, int nr, cpprob::engine::run_args * a);
    cudaError_t (*launch_rows)(cudaStream_t s, int grid, cpprob::engine::run_args * a);
    cudaError_t (*launch_replay)(cudaStream_t s, int grid, const double * obs, int n_obs, const double * real_rows,
                                 const int * int_rows, unsigned long long stride, unsigned long long n, double * logw_out);
    // resident CTAs per SM: which = 0 fused(nr=1), 1 fused(nr=2), 2 fused(nr=4), 3 rows
    int (*occupancy)(int which, int n_obs);
    // staged kernel (staged_kernels.cuh): warps per CTA its staging areas allow for this trace structure (0: does not
    // fit, use the row path), and the launch (one CTA of `warps` warps per grid entry; a->n_real/n_int/hist_* set)
    int (*staged_warps)(int n_obs, int n_real, int n_int, int bins);
    cudaError_t (*launch_staged)(cudaStream_t s, int grid, int warps, cpprob::engine::run_args * a);
    // > 0: the model declares that its integral predicts lie in [0, int_states) (Model::int_predict_states)
    int int_states;
};
}

namespace cpprob {
namespace engine {

// Dynamic shared memory of a model's kernels: the ziggurat table, if the model draws normals.  More than the 48 KB a
// kernel may use without asking, hence the attribute (set at every launch: it is per device, and cheap).
// plus the model's per-launch table (particle.hpp "Per-launch model tables"), when it fits kScratchBudget
constexpr unsigned kScratchBudget = 96u * 1024u;
template<class Model>
inline int model_scratch_doubles(int n_obs)
{
    if (!model_scratch<Model>::present) return 0;
    const int n = model_scratch<Model>::doubles(n_obs);
    return (n > 0 && static_cast<unsigned>(n) * sizeof(double) <= kScratchBudget) ? ((n + 1) & ~1) : 0;     // even: keeps 16-byte alignment
}
template<class Model>
inline unsigned model_smem(int n_obs)
{
    return zig_smem_doubles<Model>() * static_cast<unsigned>(sizeof(double)) + static_cast<unsigned>(model_scratch_doubles<Model>(n_obs)) * static_cast<unsigned>(sizeof(double));
}

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
    static cudaError_t pilot(cudaStream_t s, const philox_keys * keys, const double * obs, int n_obs, int n_pilot, double * out, const pilot_finalize * fin)
    {
        if (cudaError_t err = allow_smem(k_pilot<Model>, model_smem<Model>(n_obs))) return err;
        const pilot_finalize f = fin ? *fin : pilot_finalize{nullptr, 0, 0.0};
        k_pilot<Model><<<(n_pilot + 511) / 512, kBlock, model_smem<Model>(n_obs), s>>>(*keys, obs, n_obs, n_pilot, model_scratch_doubles<Model>(n_obs), out, f);
        return cudaGetLastError();
    }
    static cudaError_t fused(cudaStream_t s, int grid, int nr, run_args * a)
    {
        const unsigned smem = model_smem<Model>(a->n_obs);
        a->scratch_doubles = model_scratch_doubles<Model>(a->n_obs);
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
    static cudaError_t rows(cudaStream_t s, int grid, run_args * a)
    {
        a->scratch_doubles = model_scratch_doubles<Model>(a->n_obs);
        if (cudaError_t err = allow_smem(k_sis_rows<Model>, model_smem<Model>(a->n_obs))) return err;
        k_sis_rows<Model><<<grid, kBlock, model_smem<Model>(a->n_obs), s>>>(*a);
        return cudaGetLastError();
    }
    static cudaError_t replay(cudaStream_t s, int grid, const double * obs, int n_obs, const double * real_rows,
                              const int * int_rows, unsigned long long stride, unsigned long long n, double * logw_out)
    {
        // only the model table: replayed traces draw nothing, so the ziggurat part stays unused (but keeps its offset)
        if (cudaError_t err = allow_smem(k_replay<Model>, model_smem<Model>(n_obs))) return err;
        k_replay<Model><<<grid, kBlock, model_smem<Model>(n_obs), s>>>(obs, n_obs, model_scratch_doubles<Model>(n_obs), real_rows, int_rows, stride, n, logw_out);
        return cudaGetLastError();
    }
    // warps per CTA of the staged kernel: as many as the per-warp staging areas leave room for next to the tables
    static int staged_warps(int n_obs, int n_real, int n_int, int bins)
    {
        if (bins > kStagedMaxBins || (n_int > 0 && bins <= 0)) return 0;
        const unsigned fixed = model_smem<Model>(n_obs);
        const unsigned per_warp = make_stage_layout(n_real, n_int, bins, staged_packed<Model>()).bytes;
        if (fixed + 4u * per_warp > kSmemBudget) return 0;
        const unsigned w = (kSmemBudget - fixed) / per_warp;
        constexpr unsigned max_warps = staged_threads<Model>() / 32u;
        return static_cast<int>(w > max_warps ? max_warps : w);
    }
    static cudaError_t staged(cudaStream_t s, int grid, int warps, run_args * a)
    {
        a->scratch_doubles = model_scratch_doubles<Model>(a->n_obs);
        a->stage_base = model_smem<Model>(a->n_obs);
        const unsigned smem = a->stage_base + static_cast<unsigned>(warps) * make_stage_layout(a->n_real, a->n_int, a->hist_bins, staged_packed<Model>()).bytes;
        if (staged_threads<Model>() > kStagedFewThreads && warps * 32 <= kStagedFewThreads) {
            if (cudaError_t err = allow_smem(k_sis_staged<Model, kStagedFewThreads>, smem)) return err;
            k_sis_staged<Model, kStagedFewThreads><<<grid, warps * 32, smem, s>>>(*a);
        } else {
            if (cudaError_t err = allow_smem(k_sis_staged<Model, staged_threads<Model>()>, smem)) return err;
            k_sis_staged<Model, staged_threads<Model>()><<<grid, warps * 32, smem, s>>>(*a);
        }
        return cudaGetLastError();
    }
    static int occupancy(int which, int n_obs)
    {
        int n = 0;
        cudaError_t err;
        const unsigned smem = model_smem<Model>(n_obs);
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
            &probe, &pilot, &fused, &rows, &replay, &occupancy, &staged_warps, &staged, model_int_states<Model>::value};
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
