// cpprob-b200: array front-ends of the device distribution layer (log-pdfs, samplers, Philox, fp64
// elementary functions).  They back cpprob_sis_logpdf / _sample / _philox / _dmath, which the parity
// tests use to pin the layer against the reference's own known-answer grids
// (/root/reference tests/cpprob/logpdf.cpp:23-35,61-78) and against scipy.
#ifndef CPPROB_B200_DIST_KERNELS_CUH
#define CPPROB_B200_DIST_KERNELS_CUH

#include "sis_kernels.cuh"

namespace cpprob {
namespace engine {

struct dist_params { double p[8]; int n; };

template<class F>
__global__ void __launch_bounds__(kBlock) k_map(unsigned long long n, F f)
{
    const unsigned zig_base = f.prepare();
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * kBlock) {
        f(i, zig_base);
    }
}

struct logpdf_op {
    __device__ unsigned prepare() const { return 0; }
    int kind;
    dist_params q;
    const double * x;
    double * out;
    __device__ void operator()(unsigned long long i, unsigned zig_base) const
    {
        const double xi = x[i];
        double r = 0.0;
        switch (kind) {
        case CPPROB_SIS_DIST_NORMAL: {
            normal_distribution<> d(q.p[0], q.p[1]);
            r = logpdf<normal_distribution<>>()(d, xi);
        } break;
        case CPPROB_SIS_DIST_UNIFORM_REAL: {
            uniform_real_distribution<> d(q.p[0], q.p[1]);
            r = logpdf<uniform_real_distribution<>>()(d, xi);
        } break;
        case CPPROB_SIS_DIST_UNIFORM_SMALLINT: {
            uniform_smallint<long long> d(static_cast<long long>(q.p[0]), static_cast<long long>(q.p[1]));
            r = logpdf<uniform_smallint<long long>>()(d, static_cast<long long>(xi));
        } break;
        case CPPROB_SIS_DIST_DISCRETE: {
            discrete_distribution<long long, double, 8> d(q.p, q.p + q.n);
            r = logpdf<discrete_distribution<long long, double, 8>>()(d, static_cast<long long>(xi));
        } break;
        case CPPROB_SIS_DIST_POISSON: {
            poisson_distribution<long long, double> d(q.p[0]);
            r = logpdf<poisson_distribution<long long, double>>()(d, static_cast<long long>(xi));
        } break;
        case CPPROB_SIS_DIST_GAMMA: {
            gamma_distribution<> d(q.p[0], q.p[1]);
            r = logpdf<gamma_distribution<>>()(d, xi);
        } break;
        case CPPROB_SIS_DIST_BETA: {
            beta_distribution<> d(q.p[0], q.p[1]);
            r = logpdf<beta_distribution<>>()(d, xi);
        } break;
        default: r = 0.0 / 0.0;
        }
        out[i] = r;
    }
};

struct sample_op {
    __device__ unsigned prepare() const { return zig::load_shared(); }   // the normal / gamma samplers read the ziggurat tables
    int kind;
    dist_params q;
    unsigned long long seed, first;
    double * out;
    __device__ void operator()(unsigned long long i, unsigned zig_base) const
    {
        const philox_keys keys(seed);
        philox_stream rng(keys, first + i, zig_base);
        double r = 0.0;
        switch (kind) {
        case CPPROB_SIS_DIST_NORMAL: r = normal_distribution<>(q.p[0], q.p[1])(rng); break;
        case CPPROB_SIS_DIST_UNIFORM_REAL: r = uniform_real_distribution<>(q.p[0], q.p[1])(rng); break;
        case CPPROB_SIS_DIST_UNIFORM_SMALLINT:
            r = static_cast<double>(uniform_smallint<long long>(static_cast<long long>(q.p[0]), static_cast<long long>(q.p[1]))(rng));
            break;
        case CPPROB_SIS_DIST_DISCRETE:
            r = static_cast<double>(discrete_distribution<long long, double, 8>(q.p, q.p + q.n)(rng));
            break;
        case CPPROB_SIS_DIST_POISSON: r = static_cast<double>(poisson_distribution<long long, double>(q.p[0])(rng)); break;
        case CPPROB_SIS_DIST_GAMMA: r = gamma_distribution<>(q.p[0], q.p[1])(rng); break;
        case CPPROB_SIS_DIST_BETA: r = beta_distribution<>(q.p[0], q.p[1])(rng); break;
        default: r = 0.0 / 0.0;
        }
        out[i] = r;
    }
};

struct philox_op {
    __device__ unsigned prepare() const { return 0; }
    const unsigned * ctr;
    const unsigned * key;
    unsigned * out;
    __device__ void operator()(unsigned long long i, unsigned zig_base) const
    {
        unsigned o0, o1, o2, o3;
        const philox_keys keys(key[2 * i], key[2 * i + 1]);
        philox4x32::block(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], keys, o0, o1, o2, o3);
        out[4 * i] = o0; out[4 * i + 1] = o1; out[4 * i + 2] = o2; out[4 * i + 3] = o3;
    }
};

struct dmath_op {
    __device__ unsigned prepare() const { return fn == 9 ? dm::exp2_table_load() : 0u; }   // the slot carries the exp table base here
    int fn;
    const double * x;
    double * out;
    __device__ void operator()(unsigned long long i, unsigned zig_base) const
    {
        double r, s, c;
        switch (fn) {
        case 0: r = dm::log_unit(x[i]); break;
        case 1: r = dm::exp_weight(x[i]); break;
        case 2: dm::sincos_2pi(x[i], s, c); r = s; break;
        case 3: dm::sincos_2pi(x[i], s, c); r = c; break;
        case 4: r = dm::sqrt_pos(x[i]); break;
        case 6: r = dm::log(x[i]); break;
        case 7: r = dm::cos_2pi(x[i]); break;
        case 8: r = dm::sin_2pi(x[i]); break;
        case 9: r = dm::exp_weight_tab(x[i], zig_base); break;
        case 5: {
            const double rad = dm::sqrt_pos(-2.0 * dm::log_unit(x[2 * i]));
            dm::sincos_2pi(x[2 * i + 1], s, c);
            r = rad * c;
        } break;
        default: r = 0.0 / 0.0;
        }
        out[i] = r;
    }
};

}  // namespace engine
}  // namespace cpprob
#endif  // CPPROB_B200_DIST_KERNELS_CUH
