// cpprob-b200: the SIS kernels (sm_100a).  One particle = one thread-iteration; all particle state
// (Philox key/counter, log_w, staged predict values) is register-resident.
//
// What the reference does per particle (/root/reference: include/cpprob/cpprob.hpp:194-201):
//     start_trace(); call_f_tuple(f, observes); finish_trace();
// with finish_trace appending a text record to three files (src/cpprob/state.cpp:193-202), and a
// later StatsPrinter pass re-reading the text to form self-normalised estimators
// (include/cpprob/postprocess/stats_printer.hpp:88-120, empirical_distribution.hpp:52-81,117-143).
// Here the model body, the weight exp(log_w - m_ref) and the estimator sums are one kernel.
//
// Determinism: particles are grouped in fixed CHUNKs of 2^15 consecutive global indices with a fixed
// (virtual thread, warp slot) -> particle map.  A chunk's partial sums are always formed by the same tree
// (per-thread sums, warp shuffles, the 8 warp slots in slot order), so they are bit-identical whatever
// the grid size, schedule (work units are handed out through an atomic counter) or GPU count.  Chunk
// partials are folded into super-chunk rows in chunk order (k_fold_rows) and those are merged in row order
// by k_merge_columns.  All weights are taken relative to one run-wide reference m_ref (max log_w of
// a pilot over global particles [0, 4096), identical on every rank), so merging is plain addition.
#ifndef CPPROB_B200_SIS_KERNELS_CUH
#define CPPROB_B200_SIS_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "cpprob/particle.hpp"

namespace cpprob {
namespace engine {

#ifndef CPPROB_TILES_PER_TRIP
#define CPPROB_TILES_PER_TRIP 2             // stream tiles (pairs of particles per thread) per loop trip
#endif
// The fused kernel's warps are autonomous, so its CTA size is only a register budget: one CTA of 768 threads per
// SM = 32 warps at 64 registers and ONE copy of the 64 KB ziggurat table (measured best of 256x2, 256x3,
// 512 ... 1024: profiles/r01_notes.md).  Kernels that stage more predicts keep 256 threads.
#ifndef CPPROB_FUSED_MIN_BLOCKS
#define CPPROB_FUSED_MIN_BLOCKS 1           // resident CTAs per SM the one-predict fused kernel is compiled for
#endif
#ifndef CPPROB_FUSED_THREADS
#define CPPROB_FUSED_THREADS 1024           // threads per CTA of the one-predict fused kernel (any multiple of 32)
#endif

constexpr int kBlock = 256;                 // threads per CTA
constexpr int kWarps = kBlock / 32;
constexpr int fused_block(int nr) { return nr == 1 ? CPPROB_FUSED_THREADS : kBlock; }
constexpr unsigned kChunk = 1u << 15;       // particles per deterministic reduction unit (fused kernel, shard granularity)
constexpr unsigned kSubChunk = 1u << 12;    // particles per partial row on the row (SoA) path; kChunk / kSubChunk rows per chunk
constexpr int kBaseCols = 8;                // partial columns every run has (see col:: below)
constexpr int kPilot = 4096;                // pilot particles (global indices [0, kPilot))
constexpr int kPilotTiles = kPilot / 512;   // one CTA each
constexpr int kMomTile = 8;                 // real rows per k_row_moments CTA
constexpr int kMaxFusedReal = 4;            // register-staged predict slots in the fused kernel

namespace col {
enum : int { max_lw = 0, s0 = 1, s00 = 2, n_neginf = 3, neg_imin = 4, imax = 5, int_oor = 6, n_nan = 7 };
}
// columns merged with max (all others with +)
constexpr unsigned long long kMaxColsMask = (1ull << col::max_lw) | (1ull << col::neg_imin) | (1ull << col::imax);

static_assert(kBlock == static_cast<int>(kPairStride), "one CTA row of threads per half stream tile");
static_assert(kChunk % (2 * kPairStride) == 0, "chunks hold whole stream tiles");

struct int_extra;

struct run_args {
    philox_keys keys;                    // expanded Philox round keys of the run's seed
    unsigned long long first_particle;   // global index of this launch's first particle (multiple of kChunk)
    unsigned long long n_particles;      // particles in this launch
    unsigned n_chunks;                   // ceil(n_particles / kChunk)
    int n_obs;
    const double * obs;                  // device
    const double * m_ref;                // device scalar
    unsigned * chunk_counter;            // device, zeroed before the launch
    double * partials;                   // [n_chunks][n_cols]
    double * warp_partials;              // fused kernel: [n_chunks * 8][kBaseCols + 2 NR] scratch, folded into partials
    int n_cols;
    // SoA trace rows (row kernels only); column index = particle index within the launch
    double * real_rows;                  // [n_real][row_stride]
    int * int_rows;                      // [n_int][row_stride]
    double * logw;                       // [row_stride]
    double * w;                          // [row_stride]
    unsigned long long row_stride;
    // histogram window for int predicts (from the pilot)
    long long hist_lo;
    int hist_bins;
    // row path: particles per partial row (sub-chunk) and the per-sub-chunk int bookkeeping
    unsigned chunk;
    struct int_extra * int_extras;       // [n_sub_chunks] or nullptr when the model has no int predicts
    // model table (Model::fill_scratch) in shared memory: its size in doubles, 0 = none (absent / beyond the budget)
    int scratch_doubles;
    // staged kernel: the trace structure and the byte offset of the per-warp staging areas in dynamic shared memory
    int n_real, n_int;
    unsigned stage_base;
};

// ------------------------------------------------------------------------------------------------
// CTA-wide reduction of N doubles with a fixed tree.  Slot i uses max if bit i of max_mask is set.
// Result valid in threads [0, N) of the CTA (one value each).
// ------------------------------------------------------------------------------------------------
template<int N>
__device__ __forceinline__ double block_reduce(double (&v)[N], unsigned long long max_mask, double * smem /*[kWarps][N]*/)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x = v[i];
        const bool is_max = (max_mask >> i) & 1ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, off);
            x = is_max ? fmax(x, y) : x + y;
        }
        if (lane == 0) smem[warp * N + i] = x;
    }
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < N) {
        const bool is_max = (max_mask >> threadIdx.x) & 1ull;
        r = smem[threadIdx.x];
#pragma unroll
        for (int wi = 1; wi < kWarps; ++wi) {
            const double y = smem[wi * N + threadIdx.x];
            r = is_max ? fmax(r, y) : r + y;
        }
    }
    __syncthreads();
    return r;
}

// special-value tests on the bit pattern (ALU pipe) instead of DSETP (FP64 pipe)
__device__ __forceinline__ bool is_finite(double x)
{
    return (static_cast<unsigned>(__double2hiint(x)) & 0x7FF00000u) != 0x7FF00000u;
}
__device__ __forceinline__ bool is_neg_inf(double x)
{
    return __double2hiint(x) == static_cast<int>(0xFFF00000u) && __double2loint(x) == 0;
}
__device__ __forceinline__ bool is_nan(double x)
{
    const unsigned hi = static_cast<unsigned>(__double2hiint(x)) & 0x7FFFFFFFu;
    return hi > 0x7FF00000u || (hi == 0x7FF00000u && __double2loint(x) != 0);
}

// Runs body(rng, i) for the particles of [0, n_here) owned by virtual thread `vt` (0..255): in every tile of 512
// the thread owns local indices vt and vt + 256, which share one Philox stream (random/philox.hpp).
// `global_base` is the global index of local particle 0 (a multiple of 512).  vt is threadIdx.x where a CTA
// works on the range together, and warp_slot * 32 + lane in the warp-autonomous fused kernel.
template<class Body>
__device__ __forceinline__ void for_each_owned_particle(const philox_keys & keys, unsigned zig_base, unsigned vt,
                                                        unsigned long long global_base, unsigned n_here, Body && body)
{
    constexpr unsigned kTile = 2 * kPairStride;
    const unsigned long long stream0 = stream_of_particle(global_base) + vt;
    const unsigned full_tiles = n_here / kTile;
    unsigned tile = 0;
    // two whole tiles per trip: four independent particle bodies in one straight-line block, so that the
    // scheduler can overlap the integer (Philox) phase of one with the FP64 chains of another and the
    // polynomial constants are fetched once for all four
#if CPPROB_TILES_PER_TRIP == 2
    for (; tile + 2 <= full_tiles; tile += 2) {
        philox_stream r0(keys, stream0 + static_cast<unsigned long long>(tile) * kPairStride, zig_base);
        philox_stream r1(keys, stream0 + static_cast<unsigned long long>(tile + 1) * kPairStride, zig_base);
        const unsigned i0 = tile * kTile + vt;
        body(r0, i0);
        body(r1, i0 + kTile);
        body(r0, i0 + kPairStride);
        body(r1, i0 + kTile + kPairStride);
    }
#else
    for (; tile < full_tiles; ++tile) {
        philox_stream r0(keys, stream0 + static_cast<unsigned long long>(tile) * kPairStride, zig_base);
        const unsigned i0 = tile * kTile + vt;
        body(r0, i0);
        body(r0, i0 + kPairStride);
    }
#endif
    for (; tile * kTile < n_here; ++tile) {
        const unsigned ia = tile * kTile + vt;
        if (ia < n_here) {
            philox_stream rng(keys, stream0 + static_cast<unsigned long long>(tile) * kPairStride, zig_base);
            body(rng, ia);
            const unsigned ib = ia + kPairStride;
            if (ib < n_here) body(rng, ib);
        }
    }
}

// A model that never draws a normal (directly or through gamma / beta) may say so with
// `static constexpr bool draws_normals = false;`: its kernels then skip the 64 KB shared-memory ziggurat table,
// which would otherwise cap their resident CTAs.  Absent = true (always safe).
template<class Model, class = void>
struct model_draws_normals : std::true_type {};
template<class Model>
struct model_draws_normals<Model, decltype(void(Model::draws_normals))> : std::integral_constant<bool, Model::draws_normals> {};

template<class Model>
__device__ __forceinline__ unsigned zig_prepare()
{
    if (model_draws_normals<Model>::value) return zig::load_shared();
    return 0u;
}

// Dynamic shared memory of a model kernel: [ziggurat table, if the model draws normals][model table][kernel-specific].
// The ziggurat part is rounded up to 16 bytes so that what follows can be read with 128-bit loads.
template<class Model>
constexpr unsigned zig_smem_doubles() { return model_draws_normals<Model>::value ? ((zig::N + 1 + 1) & ~1u) : 0u; }

// Fills the model's per-launch table (particle.hpp, "Per-launch model tables") cooperatively, once per CTA.
// Returns nullptr when the model has none or the launch was given no room for it (scratch_doubles == 0).
template<class Model>
__device__ __forceinline__ const double * model_scratch_prepare(const double * __restrict__ obs, int n_obs, int scratch_doubles)
{
    if (!model_scratch<Model>::present || scratch_doubles <= 0) return nullptr;
    extern __shared__ double cpprob_zig_shared[];
    double * const t = cpprob_zig_shared + zig_smem_doubles<Model>();
    model_scratch<Model>::fill(t, obs, n_obs, static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x));
    __syncthreads();
    return t;
}

// Observations of scalar-argument models are read once into registers; array models read theirs
// through the read-only path as they go (the same address in every lane: one L1 hit per warp).
template<class Model, bool Scalars = (Model::n_scalar_obs >= 0)>
struct obs_cache {
    double v[Model::n_scalar_obs > 0 ? Model::n_scalar_obs : 1];
    __device__ __forceinline__ obs_cache(const double * __restrict__ obs, int)
    {
#pragma unroll
        for (int i = 0; i < Model::n_scalar_obs; ++i) v[i] = __ldg(obs + i);
    }
    __device__ __forceinline__ const double * data() const { return v; }
};
template<class Model>
struct obs_cache<Model, false> {
    const double * p;
    __device__ __forceinline__ obs_cache(const double * __restrict__ obs, int) : p(obs) {}
    __device__ __forceinline__ const double * data() const { return p; }
};

// ------------------------------------------------------------------------------------------------
// Policies (see include/cpprob/particle.hpp)
// ------------------------------------------------------------------------------------------------

// Pilot / dry run: draw from the prior, drop predicts, remember the int range.
struct null_policy {
    long long imin, imax;
    __device__ __forceinline__ null_policy() : imin(0x7fffffffffffffffLL), imax(-0x7fffffffffffffffLL - 1) {}
    template<class D>
    __device__ __forceinline__ typename D::result_type sample(const D & d, philox_stream & rng) { return d(rng); }
    template<class S> __device__ __forceinline__ void predict_int(long long x, const S &)
    {
        imin = x < imin ? x : imin;
        imax = x > imax ? x : imax;
    }
    template<class S> __device__ __forceinline__ void predict_real(double, const S &) {}
    template<class S> __device__ __forceinline__ void begin_vector(int, const S &) {}
};

// Fused kernel: up to NR real predicts staged in registers.  Lenient = the fast pass, whose units are recomputed
// with exact special-case semantics whenever a non-finite log-weight appears (see logpdf<normal>::finite_case).
template<int NR, bool Lenient = false>
struct reg_policy {
    static constexpr bool lenient_logpdf = Lenient;
    static constexpr bool first_observe_stores = true;      // see particle::observe
    double v[NR];
    int k;
    __device__ __forceinline__ reg_policy() : k(0)
    {
#pragma unroll
        for (int i = 0; i < NR; ++i) v[i] = 0.0;
    }
    template<class D>
    __device__ __forceinline__ typename D::result_type sample(const D & d, philox_stream & rng) { return d(rng); }
    template<class S> __device__ __forceinline__ void predict_int(long long, const S &) {}
    template<class S> __device__ __forceinline__ void predict_real(double x, const S &)
    {
#pragma unroll
        for (int i = 0; i < NR; ++i) if (k == i) v[i] = x;
        ++k;
    }
    template<class S> __device__ __forceinline__ void begin_vector(int, const S &) {}
};

// Integral predicts are recorded as 32-bit values (StatsPrinter parses the .int file as `int`, stats_printer.hpp:83).
// A wider value that does not fit is never truncated silently: it drives the tracked range to the whole of int32, which
// no histogram window covers, so the run ends with CPPROB_SIS_ERANGE.
template<class T>
__device__ __forceinline__ int narrow_int(T x, int & imin, int & imax)
{
    if (sizeof(T) > sizeof(int)) {
        if (static_cast<T>(static_cast<int>(x)) != x) {
            imin = static_cast<int>(0x80000000u);
            imax = 0x7fffffff;
        }
    }
    return static_cast<int>(x);
}

// Row kernel: address-major SoA rows in HBM; consecutive threads own consecutive columns, so each
// predict statement is one fully coalesced store per warp.
struct row_policy {
    double * real_col;
    int * int_col;
    unsigned long long stride;
    int imin, imax;                       // of the stored (int32) values; k_row_base checks them against the window
    __device__ __forceinline__ row_policy(double * rc, int * ic, unsigned long long s)
        : real_col(rc), int_col(ic), stride(s), imin(0x7fffffff), imax(static_cast<int>(0x80000000u)) {}
    template<class D>
    __device__ __forceinline__ typename D::result_type sample(const D & d, philox_stream & rng) { return d(rng); }
    template<class T, class S> __device__ __forceinline__ void predict_int(T x, const S &)
    {
        const int xi = narrow_int(x, imin, imax);
        *int_col = xi;
        int_col += stride;
        imin = min(imin, xi);
        imax = max(imax, xi);
    }
    template<class S> __device__ __forceinline__ void predict_real(double x, const S &)
    {
        *real_col = x;
        real_col += stride;
    }
    template<class S> __device__ __forceinline__ void begin_vector(int, const S &) {}
};

// Replay: the k-th sample statement of a kind returns the k-th recorded predict of that kind.
// Valid for models that predict every sampled value in program order (all three target models;
// SURVEY.md §8c "Replay check feasibility").
struct replay_policy {
    const double * real_col;
    const int * int_col;
    unsigned long long stride;
    __device__ __forceinline__ replay_policy(const double * rc, const int * ic, unsigned long long s)
        : real_col(rc), int_col(ic), stride(s) {}
    template<class D>
    __device__ __forceinline__ typename D::result_type sample(const D &, philox_stream &)
    {
        return take_any(static_cast<typename D::result_type *>(nullptr));
    }
    template<class S> __device__ __forceinline__ void predict_int(long long, const S &) {}
    template<class S> __device__ __forceinline__ void predict_real(double, const S &) {}
    template<class S> __device__ __forceinline__ void begin_vector(int, const S &) {}
    // one recorded real value (used by vector-valued distributions, component by component)
    __device__ __forceinline__ double take_real()
    {
        const double x = __ldg(real_col);
        real_col += stride;
        return x;
    }

private:
    template<class R>
    __device__ __forceinline__ R take_any(R * tag)
    {
        return take(std::integral_constant<bool, std::is_integral<R>::value>(), tag);
    }
    template<class T, int N>
    __device__ __forceinline__ vecn<T, N> take_any(vecn<T, N> *)
    {
        vecn<T, N> r;
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = static_cast<T>(take_real());
        return r;
    }
    template<class R>
    __device__ __forceinline__ R take(std::true_type, R *)
    {
        const int x = __ldg(int_col);
        int_col += stride;
        return static_cast<R>(x);
    }
    template<class R>
    __device__ __forceinline__ R take(std::false_type, R *)
    {
        const double x = __ldg(real_col);
        real_col += stride;
        return static_cast<R>(x);
    }
};

// ------------------------------------------------------------------------------------------------
// K_pilot: max log_w and int-predict range over global particles [0, n_pilot).  One CTA per stream
// tile (512 particles), so that long traces (hmm<1000>) do not serialise on a single SM.
// out[3*cta + {0,1,2}] = max log_w, max(-int), max(int) of that tile (-inf where there is none); the
// host takes the maximum over the kPilot/512 tiles (max is order-independent: deterministic).
// ------------------------------------------------------------------------------------------------
// With `fin.sync` (two zero-initialised words; word 1 counts the CTAs that are done and is left at zero again) the last
// CTA to finish also does what the host would: out[0] = m_ref — the override if given, else the maximum over the tiles if
// it is finite and sane, else 0 (every pilot weight was -inf / nan) — and zeroes word 0, the unit counter of the particle
// kernel that follows.  No second launch, no copy, no memset between the pilot and the particle kernel.
struct pilot_finalize {
    unsigned * sync;            // nullptr: leave the tiles' maxima for the host (run_pilot)
    int has_override;
    double override_value;
};

template<class Model>
__global__ void __launch_bounds__(kBlock) k_pilot(const __grid_constant__ philox_keys keys, const double * __restrict__ obs,
                                                  int n_obs, int n_pilot, int scratch_doubles, double * __restrict__ out,
                                                  const pilot_finalize fin)
{
    constexpr unsigned kTile = 2 * kPairStride;
    __shared__ double smem[kWarps * 3];
    const Model model{};
    const obs_cache<Model> oc(obs, n_obs);
    const unsigned zig_base = zig_prepare<Model>();
    const double * const scratch = model_scratch_prepare<Model>(obs, n_obs, scratch_doubles);
    double v[3] = {dm::neg_inf(), dm::neg_inf(), dm::neg_inf()};   // max lw, max(-imin), max(imax)
    const unsigned base = blockIdx.x * kTile;
    const unsigned n_here = static_cast<unsigned>(n_pilot) > base ? min(static_cast<unsigned>(n_pilot) - base, kTile) : 0u;
    for_each_owned_particle(keys, zig_base, threadIdx.x, static_cast<unsigned long long>(base), n_here, [&](philox_stream & rng, unsigned) {
        null_policy pol;
        particle<null_policy> p(rng, pol, scratch);
        invoke_model(model, p, oc.data(), n_obs);
        v[0] = fmax(v[0], p.log_w());
        if (pol.imin <= pol.imax) {
            v[1] = fmax(v[1], -static_cast<double>(pol.imin));
            v[2] = fmax(v[2], static_cast<double>(pol.imax));
        }
    });
    const double r = block_reduce<3>(v, 0x7ull, smem);
    if (threadIdx.x < 3) out[3 * blockIdx.x + threadIdx.x] = r;
    if (fin.sync == nullptr) return;
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(fin.sync + 1, 1u) + 1u == gridDim.x;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double mx = dm::neg_inf();
        for (unsigned t = 0; t < gridDim.x; ++t) mx = fmax(mx, __ldcg(out + 3 * t));   // written by the other CTAs
        const double m = (mx > -1.0e300 && mx < 1.0e300) ? mx : 0.0;
        out[0] = fin.has_override ? fin.override_value : m;
        fin.sync[1] = 0u;
        fin.sync[0] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// K1 k_sis_fused: model body + weight + estimator sums, no trace written.  Used when the model has
// no int predicts and at most NR real predicts per trace and no trace emission was asked for.
//
// Warp-autonomous: the work unit is (chunk c, warp slot w) = the particles that lanes [32w, 32w+32) of a
// 256-thread CTA would own in chunk c (4096 particles).  Any warp of the grid takes any unit from the atomic
// counter, sums it with a fixed shuffle tree and writes NV values to warp_partials[c*8 + w][*]; after the one
// barrier of the table load no warp ever waits for another (the ziggurat's rare slow draws make warps finish
// at different times).  k_fold_units (reduce_kernels.cuh) then adds the 8 rows of a chunk in row order, which is the sum
// the CTA-wide tree used to form: chunk partials are bit-identical for any grid size / schedule / GPU count.
// Partial columns: kBaseCols then (S1, S2) per real predict slot.
// ------------------------------------------------------------------------------------------------
constexpr int kSlotsPerChunk = kBlock / 32;          // warp slots of one chunk
// The fused kernel can cut every (chunk, warp slot) into kFusedParts units of consecutive stream tiles.  Measured with 2
// (2048-particle units, profiles/r02_notes.md): the end of a short run does not get shorter (1.25e8 particles: 0.425 vs
// 0.427 ms) and a long one pays the second reduction (1e9: 3.080 vs 3.037 ms), so the shipped value is 1.
#ifndef CPPROB_FUSED_PARTS
#define CPPROB_FUSED_PARTS 1
#endif
constexpr int kFusedParts = CPPROB_FUSED_PARTS;
constexpr int kFusedRowsPerChunk = kSlotsPerChunk * kFusedParts;   // warp_partials rows of one chunk, folded in row order
constexpr unsigned kFusedPart = kChunk / kFusedParts;              // particles a part spans (all eight slots together)
static_assert(kFusedPart % (4 * kPairStride) == 0, "a part holds whole pairs of stream tiles");
static_assert(kFusedParts >= 1 && kFusedParts <= 4, "the unit counter is 32 bits wide (run_shard_impl's limit assumes <= 4 parts)");

// ------------------------------------------------------------------------------------------------
// Warp sum of C columns at once ("transpose-reduce").  Every lane brings v[0..C); at the xor-16 stage the lanes with
// bit 4 clear keep the lower half of the columns and hand the upper half to their partner (and vice versa), so each
// exchange finishes one column pair instead of one column; once a lane is down to one column the remaining stages are
// the plain butterfly.  Every column is summed by exactly the tree the plain xor-butterfly forms (same pairs, IEEE
// addition commutes), so the result has the same bits; it just costs 6 exchanges instead of 20 for four columns.
// Afterwards v[0] of lane L holds the total of column transpose_col<C>(L) (or nothing: -1).
// ------------------------------------------------------------------------------------------------
template<int P, int OFF>
__device__ __forceinline__ void warp_transpose_sum(double * v, unsigned lane)
{
    if constexpr (OFF >= 1) {
        if constexpr (P == 1) {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
            warp_transpose_sum<1, OFF / 2>(v, lane);
        } else {
            constexpr int H = (P + 1) / 2;
            const bool hi = (lane & OFF) != 0u;
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const double upper = (H + j) < P ? v[H + j] : 0.0;
                const double send = hi ? v[j] : upper;
                const double keep = hi ? upper : v[j];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            warp_transpose_sum<H, OFF / 2>(v, lane);
        }
    }
}

// the column whose total lane `lane` holds after warp_transpose_sum<C, 16>, or -1 (a lane that duplicates another's
// total, or whose slot was padding)
template<int C>
__device__ __forceinline__ int transpose_col(unsigned lane)
{
    int base = 0, valid = C, p = C;
    bool owner = true;
#pragma unroll
    for (unsigned off = 16; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0u;
        if (p == 1) {
            owner = owner && !hi;              // plain stage: both partners end up with the same total
        } else {
            const int h = (p + 1) / 2;
            if (hi) { base += h; valid -= h; } else { valid = valid < h ? valid : h; }
            p = h;
        }
    }
    return owner && valid >= 1 ? base : -1;
}

template<class Model, int NR>
__global__ void __launch_bounds__(fused_block(NR), NR == 1 ? CPPROB_FUSED_MIN_BLOCKS : 1) k_sis_fused(const __grid_constant__ run_args a)
{
    constexpr int NV = kBaseCols + 2 * NR;
    const Model model{};
    const double m_ref = *a.m_ref;
    const obs_cache<Model> oc(a.obs, a.n_obs);
    const unsigned zig_base = zig_prepare<Model>();
    const double * const scratch = model_scratch_prepare<Model>(a.obs, a.n_obs, a.scratch_doubles);
    const unsigned exp_tab = dm::exp2_table_load();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_units = a.n_chunks * kFusedRowsPerChunk;
    constexpr int NS = 2 + 2 * NR;                                  // columns that are sums of doubles: S0, S00, (S1, S2) per predict
    const int sum_col = transpose_col<NS>(lane);                    // which of them this lane writes (-1: none)
    const int out_col = sum_col < 0 ? -1 : (sum_col == 0 ? static_cast<int>(col::s0) : (sum_col == 1 ? static_cast<int>(col::s00) : kBaseCols + sum_col - 2));

    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(a.chunk_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        // unit = (chunk * kFusedParts + part) * 8 + warp slot
        const unsigned vt = (unit % kSlotsPerChunk) * 32u + lane;
        const unsigned long long base = static_cast<unsigned long long>(unit / kSlotsPerChunk) * kFusedPart;
        const unsigned long long left = a.n_particles > base ? a.n_particles - base : 0ull;
        const unsigned n_here = left < kFusedPart ? static_cast<unsigned>(left) : kFusedPart;

        double max_lw, s0, s00;
        unsigned n_neginf, n_nan;
        double s1[NR], s2[NR];
        // Fast pass: weights by exp_weight_tab, no per-particle special-case handling; only the
        // lowest exp argument is tracked, on its high word (one ALU instruction per particle).  A non-finite log_w needs no
        // tracking: -inf, +inf and NaN all come out of exp_weight_tab as NaN (inf - inf in its range
        // reduction) and poison the unit's weight sum.  If any particle of the unit had a non-finite
        // log_w or a weight below the normal range, the whole unit is recomputed by the careful pass.
        // The choice depends only on the unit's own data, so results stay deterministic.
        unsigned arg_key = 0;
        auto reset = [&] {
            max_lw = dm::neg_inf(); s0 = 0.0; s00 = 0.0; n_neginf = 0; n_nan = 0;
#pragma unroll
            for (int j = 0; j < NR; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
        };
        auto accumulate = [&](double lw, double w, const double (&pv)[NR]) {
            max_lw = lw > max_lw ? lw : max_lw;
            s0 += w;
            s00 = fma(w, w, s00);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const double wx = w * pv[j];
                s1[j] += wx;
                s2[j] = fma(wx, pv[j], s2[j]);
            }
        };
        reset();
        for_each_owned_particle(a.keys, zig_base, vt, a.first_particle + base, n_here, [&](philox_stream & rng, unsigned) {
            reg_policy<NR, true> pol;
            particle<reg_policy<NR, true>> p(rng, pol, scratch);
            invoke_model(model, p, oc.data(), a.n_obs);
            const double lw = p.log_w();
            const double arg = lw - m_ref;
            const double w = dm::exp_weight_tab(arg, exp_tab);
            arg_key = max(arg_key, dm::exp_arg_key(arg));
            accumulate(lw, w, pol.v);
        });
        if (__any_sync(0xffffffffu, arg_key > dm::kExpArgKeyLimit || is_nan(s0))) {
            reset();
            for_each_owned_particle(a.keys, zig_base, vt, a.first_particle + base, n_here, [&](philox_stream & rng, unsigned) {
                reg_policy<NR> pol;
                particle<reg_policy<NR>> p(rng, pol, scratch);
                invoke_model(model, p, oc.data(), a.n_obs);
                const double lw = p.log_w();
                const double w = dm::exp_weight(lw - m_ref);
                n_neginf += is_neg_inf(lw) ? 1u : 0u;
                n_nan += is_nan(lw) ? 1u : 0u;
                accumulate(lw, w, pol.v);
            });
        }

        // The unit's row of warp_partials.  Sums of doubles: transpose-reduce (the butterfly's tree, a third of its
        // exchanges); the maximum: butterfly; the two counters: integer warp sums (exact either way); the three columns
        // only int predicts feed are constants here.
        double t[NS];
        t[0] = s0;
        t[1] = s00;
#pragma unroll
        for (int j = 0; j < NR; ++j) { t[2 + 2 * j] = s1[j]; t[3 + 2 * j] = s2[j]; }
        warp_transpose_sum<NS, 16>(t, lane);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) max_lw = fmax(max_lw, __shfl_xor_sync(0xffffffffu, max_lw, off));
        n_neginf = __reduce_add_sync(0xffffffffu, n_neginf);
        n_nan = __reduce_add_sync(0xffffffffu, n_nan);
        double * const out = a.warp_partials + static_cast<size_t>(unit) * NV;
        if (out_col >= 0) out[out_col] = t[0];
        if (lane == 1) {
            out[col::max_lw] = max_lw;
            out[col::n_neginf] = static_cast<double>(n_neginf);
            out[col::n_nan] = static_cast<double>(n_nan);
        }
        if (lane == 2) {
            out[col::neg_imin] = dm::neg_inf();
            out[col::imax] = dm::neg_inf();
            out[col::int_oor] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 k_sis_rows: model body -> SoA trace rows + log_w in HBM.  No reduction happens here, so the work
// unit is one stream tile (512 particles, two per thread) and tiles are spread over the grid with a
// plain stride: a batch of long traces (hmm<1000>: 4 KB per particle) still fills every SM even when
// it holds only a handful of reduction chunks.  The int range / out-of-window bookkeeping is folded
// into per-sub-chunk integers with order-independent integer atomics (deterministic).
// The weights, the base sums and the per-row estimator sums are formed afterwards by k_row_base,
// k_row_moments and k_row_hist, which stream the rows back (HBM/L2-bound, FP64 pipe nearly idle).
// ------------------------------------------------------------------------------------------------
struct int_extra { int vmin, vmax; };

template<class Model>
__global__ void __launch_bounds__(kBlock) k_sis_rows(const __grid_constant__ run_args a)
{
    constexpr unsigned kTile = 2 * kPairStride;
    const Model model{};
    const obs_cache<Model> oc(a.obs, a.n_obs);
    const unsigned n_tiles = static_cast<unsigned>((a.n_particles + kTile - 1) / kTile);
    const unsigned long long stream0 = stream_of_particle(a.first_particle) + threadIdx.x;
    const unsigned zig_base = zig_prepare<Model>();
    const double * const scratch = model_scratch_prepare<Model>(a.obs, a.n_obs, a.scratch_doubles);

    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long base = static_cast<unsigned long long>(tile) * kTile;
        const unsigned long long left = a.n_particles - base;
        const unsigned n_here = left < kTile ? static_cast<unsigned>(left) : kTile;
        int vmin = 0x7fffffff, vmax = static_cast<int>(0x80000000u);
        if (threadIdx.x < n_here) {
            philox_stream rng(a.keys, stream0 + static_cast<unsigned long long>(tile) * kPairStride, zig_base);
#pragma unroll 1
            for (unsigned turn = 0; turn < 2; ++turn) {
                const unsigned i = threadIdx.x + turn * kPairStride;
                if (i < n_here) {
                    const unsigned long long colidx = base + i;
                    row_policy pol(a.real_rows + colidx, a.int_rows + colidx, a.row_stride);
                    particle<row_policy> p(rng, pol, scratch);
                    invoke_model(model, p, oc.data(), a.n_obs);
                    a.logw[colidx] = p.log_w();
                    vmin = min(vmin, pol.imin);
                    vmax = max(vmax, pol.imax);
                }
            }
        }
        if (a.int_extras != nullptr) {
            // a tile never straddles a sub-chunk (sub-chunk sizes are multiples of 512)
            vmin = __reduce_min_sync(0xffffffffu, vmin);
            vmax = __reduce_max_sync(0xffffffffu, vmax);
            if ((threadIdx.x & 31) == 0 && vmin <= vmax) {
                int_extra * x = a.int_extras + base / a.chunk;
                atomicMin(&x->vmin, vmin);
                atomicMax(&x->vmax, vmax);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5 k_replay: recompute log_w from a supplied SoA trace (parity gate: <= 1e-12 relative against
// the reference/oracle log-weights).
// ------------------------------------------------------------------------------------------------
template<class Model>
__global__ void __launch_bounds__(kBlock) k_replay(const double * __restrict__ obs, int n_obs, int scratch_doubles,
                                                   const double * __restrict__ real_rows, const int * __restrict__ int_rows,
                                                   unsigned long long stride, unsigned long long n,
                                                   double * __restrict__ logw_out)
{
    const Model model{};
    const obs_cache<Model> oc(obs, n_obs);
    const double * const scratch = model_scratch_prepare<Model>(obs, n_obs, scratch_doubles);
    const philox_keys keys(0u, 0u);   // never drawn from: every sample statement returns a recorded value
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * kBlock) {
        replay_policy pol(real_rows + i, int_rows + i, stride);
        philox_stream rng(keys, i);
        particle<replay_policy> p(rng, pol, scratch);
        invoke_model(model, p, oc.data(), n_obs);
        logw_out[i] = p.log_w();
    }
}

}  // namespace engine
}  // namespace cpprob
#endif  // CPPROB_B200_SIS_KERNELS_CUH
