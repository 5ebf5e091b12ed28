// cpprob-b200: K1s k_sis_staged — model body + weight + per-(address, k) estimator sums in ONE kernel for models with
// many predicts per trace (linear_gaussian_1d<32>: 32 reals, hmm<64>: 64 ints), with no trace row in HBM.
//
// The weight of a particle is known only after its last observe (cpprob.hpp:87-89 accumulates log_w over the whole
// trace), but the sums StatsPrinter needs are per predict statement: sum_p w_p x_{p,k}, sum_p w_p x_{p,k}^2 and
// sum_p w_p [x_{p,k} == v] for every (id, k) (stats_printer.hpp:88-120, empirical_distribution.hpp:30-40,52-81).  That is a
// transposition: lanes own particles while a trace is generated, and must own rows (id, k) when the sums are formed.
//
// Each warp does it through its own staging area in shared memory:
//   1. the 32 lanes run 32 particles; every predict is one store into stage[k][lane] (one conflict-free wavefront);
//   2. each lane computes w = exp(log_w - m_ref) of its particle and stores it into wst[lane];
//   3. lane k walks row k of the stage (32 particles, 128-bit loads) with the broadcast weights and adds the round's
//      contribution to the sums of row k (and k + 32, ... for longer traces).
// No barrier is shared with another warp, no row reaches HBM, and a round costs 3 FP64 instructions + 1 shared-memory
// load per (particle, real predict), or one compare + predicated add per bin per (particle, int predict).
//
// Canonical summation order (what makes results bit-identical for any grid / GPU count, and equal to the row path):
//   work unit = (sub-chunk c of 4096 particles, warp slot s) = the 512 particles p = 512 t + 256 u + 32 s + l
//   (tile t < 8, turn u < 2, lane l) that slot s of a 256-thread CTA owns; the unit's sums are formed sequentially over
//   those particles in (t, u, l) order, per row and per bin; the 8 slots are then added in slot order
//   (k_fold_units), sub-chunks in order (k_fold_units, k_merge_columns).  Base columns (max log_w, sum w, sum w^2,
//   counts) keep the order of k_row_base: per (slot, lane) over the 16 rounds, then the xor tree over lanes, then slots.
//   The row path's reduction kernels (k_rows_moments / k_rows_hist in reduce_kernels.cuh) stage the rows they read from
//   HBM the same way and call the same round functions, so both paths produce the same bits.
#ifndef CPPROB_B200_STAGED_KERNELS_CUH
#define CPPROB_B200_STAGED_KERNELS_CUH

#include "sis_kernels.cuh"

namespace cpprob {
namespace engine {

constexpr int kStageRealStride = 34;        // doubles per staged real row: 32 lanes + 2, so lane k's 128-bit reads of row k are conflict-free
constexpr int kStageIntStride = 32;         // bytes per staged int row: one state byte per lane
constexpr int kStagedMaxBins = 64;          // widest histogram window the staged kernel takes (accumulators live in shared memory)
// CTA size the staged kernel is compiled for (it is launched with as many warps as the staging areas allow, at most
// this).  A model that draws normals carries the 64 KB ziggurat table, which leaves room for about 16 staging areas of a
// 32-predict trace: 512 threads, and the 128 registers that go with them keep the normal sampler's state out of local
// memory.  Models without it (hmm) are integer-bound and want every warp they can get: 1024 threads at 64 registers.
template<class Model>
constexpr int staged_threads() { return model_draws_normals<Model>::value ? 512 : 1024; }
constexpr unsigned kSmemBudget = 227u * 1024u;

// Histogram accumulators of one group of 32 int rows: [bins + 1][32 lanes] doubles, bin-major — the accumulator of
// (bin s, lane k) sits at s * 32 + k, so the 32 lanes always hit 32 different 8-byte bank pairs whatever their states
// are.  Slot `bins` takes the states that match no bin.  hist_acc_stride = doubles per group.
__host__ __device__ constexpr unsigned hist_acc_stride(unsigned bins) { return (bins + 1u) * 32u; }

// A model whose integral predicts all lie in [0, K) may say so: `static constexpr int int_predict_states = K;`.  The
// engine then knows the histogram window without looking (no host round trip for the pilot's int range), and for K <= 4
// the staged kernel keeps FOUR states per staged byte (2 bits each) — a 1000-step hmm trace is 8 KB per warp instead
// of 32.  A predict outside the declared range ends the run with CPPROB_SIS_ERANGE.  Absent = 0 (unknown).
template<class Model, class = void>
struct model_int_states : std::integral_constant<int, 0> {};
template<class Model>
struct model_int_states<Model, decltype(void(Model::int_predict_states))> : std::integral_constant<int, Model::int_predict_states> {};
template<class Model>
constexpr bool staged_packed() { return model_int_states<Model>::value >= 1 && model_int_states<Model>::value <= 4; }

// Histogram accumulators of a unit stay in the warp's staging area while they are small; for long traces (hmm<1000>: 24 KB
// per unit) they live in the unit's own output row in global memory (L2) and every round reads / updates / writes them
// back through a two-group scratch.  Same decision on host (launch size) and device (layout).
__host__ __device__ inline bool staged_hacc_global(int n_int, int bins)
{
    return bins > 0 && static_cast<unsigned>((n_int + 31) / 32) * hist_acc_stride(static_cast<unsigned>(bins)) * 8u > 8192u;
}

// One warp's staging area.  All offsets are multiples of 16 bytes.
struct stage_layout {
    unsigned real_off, int_off, w_off, macc_off, hacc_off, bytes;
};
__host__ __device__ inline stage_layout make_stage_layout(int n_real, int n_int, int bins, bool packed)
{
    stage_layout L;
    unsigned o = 0;
    L.real_off = o;  o += static_cast<unsigned>(n_real) * kStageRealStride * 8u;
    L.int_off = o;   o += static_cast<unsigned>(packed ? (n_int + 3) / 4 : n_int) * kStageIntStride;
    o = (o + 15u) & ~15u;
    L.w_off = o;     o += 32u * 8u;
    L.macc_off = o;  o += static_cast<unsigned>(n_real) * 16u;
    const unsigned groups = staged_hacc_global(n_int, bins) ? 2u : static_cast<unsigned>((n_int + 31) / 32);
    L.hacc_off = o;  o += groups * (bins > 0 ? hist_acc_stride(static_cast<unsigned>(bins)) : 0u) * 8u;
    L.bytes = (o + 15u) & ~15u;
    return L;
}

// ------------------------------------------------------------------------------------------------
// Round functions: the contribution of 32 staged particles to the sums of the row this lane owns.
// Explicitly rounded operations (no FMA contraction left to the compiler): the staged kernel and the row path's
// reduction kernels must produce the same bits.  Lanes beyond the end of a ragged round have w = 0 and x = 0 / an
// unmatched state staged for them, so they add exactly nothing.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void moments_round(const double * __restrict__ row, const double * __restrict__ wst, double & s1, double & s2)
{
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const double2 x = *reinterpret_cast<const double2 *>(row + j);
        const double2 w = *reinterpret_cast<const double2 *>(wst + j);
        const double wx0 = __dmul_rn(w.x, x.x);              // empirical_distribution.hpp:52-71: sum w x, sum w x^2
        s1 = __dadd_rn(s1, wx0);
        s2 = __fma_rn(wx0, x.x, s2);
        const double wx1 = __dmul_rn(w.y, x.y);
        s1 = __dadd_rn(s1, wx1);
        s2 = __fma_rn(wx1, x.y, s2);
    }
}

// States are staged as bytes (value - window start).  Lane k owns the accumulators of row k in shared memory and adds
// each particle's weight to the one its state selects: one indexed load / add / store per (particle, row), whatever the
// number of bins, and no data-dependent control flow.  (Accumulators in registers were tried first: `if (s == v) h[v] += w`
// compiles to a branch around every add, on which the lanes of a warp — different rows — diverge; a predicated add
// becomes an unconditional DADD plus two selects per (particle, bin), 15 issue slots per particle for 3 bins against 8
// here.  profiles/r02_notes.md.)  A byte at or beyond `bins` (outside the window, or the 255 of a lane beyond the end) is
// redirected to the scratch slot.  acc: this lane's column of its group's accumulators, i.e. acc[s * 32] is bin s.
// `shift` / `mask`: where the row's state sits in each staged byte — (0, 0xff) for one state per byte, (2 (k & 3), 3) for
// the packed form, whose byte row k / 4 holds rows 4 (k / 4) .. + 3.
__device__ __forceinline__ void hist_round(const unsigned char * __restrict__ row, const double * __restrict__ wst, unsigned bins, double * __restrict__ acc,
                                           unsigned shift = 0u, unsigned mask = 0xffu)
{
    const uint4 a = *reinterpret_cast<const uint4 *>(row), b = *reinterpret_cast<const uint4 *>(row + 16);
    const unsigned words[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double2 w01 = *reinterpret_cast<const double2 *>(wst + 4 * q), w23 = *reinterpret_cast<const double2 *>(wst + 4 * q + 2);
        const double w[4] = {w01.x, w01.y, w23.x, w23.y};
        const unsigned word = words[q] >> shift;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned s = min((word >> (8 * e)) & mask, bins) * 32u;
            acc[s] = __dadd_rn(acc[s], w[e]);                    // empirical_distribution.hpp:30-40: sum of w over x == v
        }
    }
}

// two rows of the same lane interleaved: their accumulator chains (load -> add -> store) are independent
__device__ __forceinline__ void hist_round2(const unsigned char * __restrict__ row0, const unsigned char * __restrict__ row1,
                                            const double * __restrict__ wst, unsigned bins, double * __restrict__ acc0, double * __restrict__ acc1,
                                            unsigned shift = 0u, unsigned mask = 0xffu)
{
    const uint4 a0 = *reinterpret_cast<const uint4 *>(row0), b0 = *reinterpret_cast<const uint4 *>(row0 + 16);
    const uint4 a1 = *reinterpret_cast<const uint4 *>(row1), b1 = *reinterpret_cast<const uint4 *>(row1 + 16);
    const unsigned w0[8] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w};
    const unsigned w1[8] = {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double2 w01 = *reinterpret_cast<const double2 *>(wst + 4 * q), w23 = *reinterpret_cast<const double2 *>(wst + 4 * q + 2);
        const double w[4] = {w01.x, w01.y, w23.x, w23.y};
        const unsigned x0 = w0[q] >> shift, x1 = w1[q] >> shift;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned s0 = min((x0 >> (8 * e)) & mask, bins) * 32u;
            const unsigned s1 = min((x1 >> (8 * e)) & mask, bins) * 32u;
            const double t0 = acc0[s0], t1 = acc1[s1];
            acc0[s0] = __dadd_rn(t0, w[e]);
            acc1[s1] = __dadd_rn(t1, w[e]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Policy: predicts go to this lane's column of the warp's staging area.
// ------------------------------------------------------------------------------------------------
template<bool Packed>
struct staged_policy {
    double * real_col;             // &stage_real[k][lane], k = next real predict
    unsigned char * int_col;       // &stage_int[k][lane] (one state per byte) or &stage_int[k / 4][lane] (Packed: four per byte)
    int lo;                        // start of the histogram window
    int imin, imax;                // range of the int predicts of this particle
    unsigned acc, sh;              // Packed: the byte being filled and the position of the next state in it
    __device__ __forceinline__ staged_policy(double * rc, unsigned char * ic, int lo_)
        : real_col(rc), int_col(ic), lo(lo_), imin(0x7fffffff), imax(static_cast<int>(0x80000000u)), acc(0u), sh(0u) {}
    template<class D>
    __device__ __forceinline__ typename D::result_type sample(const D & d, philox_stream & rng) { return d(rng); }
    template<class T, class S> __device__ __forceinline__ void predict_int(T x, const S &)
    {
        const int xi = narrow_int(x, imin, imax);
        // the low bits are enough: a value outside the window makes the run repeat with a wider one / fail (run_full)
        if (Packed) {
            acc |= (static_cast<unsigned>(xi - lo) & 3u) << sh;
            sh += 2u;
            if (sh == 8u) {
                *int_col = static_cast<unsigned char>(acc);
                int_col += kStageIntStride;
                acc = 0u;
                sh = 0u;
            }
        } else {
            *int_col = static_cast<unsigned char>(xi - lo);
            int_col += kStageIntStride;
        }
        imin = min(imin, xi);
        imax = max(imax, xi);
    }
    // after the model body: the last, partly filled byte
    __device__ __forceinline__ void finish()
    {
        if (Packed && sh != 0u) *int_col = static_cast<unsigned char>(acc);
    }
    template<class S> __device__ __forceinline__ void predict_real(double x, const S &)
    {
        *real_col = x;
        real_col += kStageRealStride;
    }
    template<class S> __device__ __forceinline__ void begin_vector(int, const S &) {}
};

// ------------------------------------------------------------------------------------------------
// K1s k_sis_staged.  a.n_chunks = number of sub-chunks (kSubChunk particles) of this launch; a.warp_partials =
// [n_chunks * 8][n_cols]; a.stage_base = byte offset of the per-warp staging areas in dynamic shared memory.
// Any warp takes any (sub-chunk, slot) unit from the atomic counter.
// ------------------------------------------------------------------------------------------------
// Threads: the CTA size the instantiation is compiled for (register budget 65536 / Threads).  Besides the model's
// default (staged_threads) a 640-thread variant serves long traces, whose staging areas leave room for <= 20 warps
// anyway: 100 registers instead of 64 keep the unit's bookkeeping out of local memory.
constexpr int kStagedFewThreads = 640;

template<class Model, int Threads>
__global__ void __launch_bounds__(Threads, 1) k_sis_staged(const __grid_constant__ run_args a)
{
    extern __shared__ double cpprob_zig_shared[];
    constexpr unsigned kTile = 2 * kPairStride;
    const Model model{};
    const double m_ref = *a.m_ref;
    const obs_cache<Model> oc(a.obs, a.n_obs);
    const unsigned zig_base = zig_prepare<Model>();
    const double * const scratch = model_scratch_prepare<Model>(a.obs, a.n_obs, a.scratch_doubles);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int n_real = a.n_real, n_int = a.n_int, bins = a.hist_bins;
    const int lo = static_cast<int>(a.hist_lo);
    constexpr bool kPacked = staged_packed<Model>();
    const bool hacc_global = staged_hacc_global(n_int, bins);      // warp-uniform
    const stage_layout L = make_stage_layout(n_real, n_int, bins, kPacked);
    const int hist0 = kBaseCols + 2 * n_real;
    char * const area = reinterpret_cast<char *>(cpprob_zig_shared) + a.stage_base + static_cast<size_t>(warp) * L.bytes;
    double * const stage_real = reinterpret_cast<double *>(area + L.real_off);
    unsigned char * const stage_int = reinterpret_cast<unsigned char *>(area + L.int_off);
    double * const wst = reinterpret_cast<double *>(area + L.w_off);
    double * const macc = reinterpret_cast<double *>(area + L.macc_off);
    double * const hacc = reinterpret_cast<double *>(area + L.hacc_off);
    const unsigned n_units = a.n_chunks * kSlotsPerChunk;
    const int n_cols = a.n_cols;

    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(a.chunk_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const unsigned c = unit / kSlotsPerChunk;
        const unsigned vt = (unit % kSlotsPerChunk) * 32u + lane;
        const unsigned long long base = static_cast<unsigned long long>(c) * kSubChunk;
        const unsigned long long left = a.n_particles - base;
        const unsigned n_here = left < kSubChunk ? static_cast<unsigned>(left) : kSubChunk;

        for (int k = static_cast<int>(lane); k < 2 * n_real; k += 32) macc[k] = 0.0;
        const int hstride = static_cast<int>(hist_acc_stride(static_cast<unsigned>(bins)));       // doubles per group of 32 rows
        double * const out = a.warp_partials + static_cast<size_t>(unit) * n_cols;
        if (hacc_global) {
            for (int k = static_cast<int>(lane); k < n_int * bins; k += 32) out[hist0 + k] = 0.0;
        } else {
            for (int k = static_cast<int>(lane); k < ((n_int + 31) / 32) * hstride; k += 32) hacc[k] = 0.0;
        }
        double max_lw = dm::neg_inf(), s0 = 0.0, s00 = 0.0;
        unsigned n_neginf = 0, n_nan = 0;
        int imin = 0x7fffffff, imax = static_cast<int>(0x80000000u);
        __syncwarp();

        const unsigned long long stream0 = stream_of_particle(a.first_particle + base) + vt;
        const unsigned n_tiles = (n_here + kTile - 1) / kTile;
        for (unsigned tile = 0; tile < n_tiles; ++tile) {
            philox_stream rng(a.keys, stream0 + static_cast<unsigned long long>(tile) * kPairStride, zig_base);
#pragma unroll 1
            for (unsigned turn = 0; turn < 2; ++turn) {
                const unsigned i = tile * kTile + turn * kPairStride + vt;
                const bool valid = i < n_here;
                if (!__any_sync(0xffffffffu, valid)) break;          // the whole round lies beyond the end
                double w = 0.0;
                if (valid) {
                    staged_policy<kPacked> pol(stage_real + lane, stage_int + lane, lo);
                    particle<staged_policy<kPacked>> p(rng, pol, scratch);
                    invoke_model(model, p, oc.data(), a.n_obs);
                    pol.finish();
                    const double lw = p.log_w();
                    w = dm::exp_weight(lw - m_ref);
                    // the base sums of k_row_base, same operations in the same order
                    max_lw = lw > max_lw ? lw : max_lw;
                    s0 += w;
                    s00 = fma(w, w, s00);
                    n_neginf += is_neg_inf(lw) ? 1u : 0u;
                    n_nan += is_nan(lw) ? 1u : 0u;
                    imin = min(imin, pol.imin);
                    imax = max(imax, pol.imax);
                } else {
                    for (int k = 0; k < n_real; ++k) stage_real[k * kStageRealStride + lane] = 0.0;
                    for (int k = 0; k < (kPacked ? (n_int + 3) / 4 : n_int); ++k) stage_int[k * kStageIntStride + lane] = 0xffu;
                }
                wst[lane] = w;
                __syncwarp();
                for (int k = static_cast<int>(lane); k < n_real; k += 32) {
                    double s1 = macc[2 * k], s2 = macc[2 * k + 1];
                    moments_round(stage_real + k * kStageRealStride, wst, s1, s2);
                    macc[2 * k] = s1;
                    macc[2 * k + 1] = s2;
                }
                {
                    // row k = 32 grp + lane: column `lane` of group grp; two groups at a time (independent chains)
                    const unsigned ubins = static_cast<unsigned>(bins);
                    const unsigned shift = kPacked ? 2u * (lane & 3u) : 0u, mask = kPacked ? 3u : 0xffu;
                    auto row_of = [&](int k) { return stage_int + (kPacked ? k >> 2 : k) * kStageIntStride; };
                    int k = static_cast<int>(lane), grp = 0;
                    if (!hacc_global) {
                        for (; k + 32 < n_int; k += 64, grp += 2) {
                            hist_round2(row_of(k), row_of(k + 32), wst, ubins, hacc + grp * hstride + lane, hacc + (grp + 1) * hstride + lane, shift, mask);
                        }
                        if (k < n_int) hist_round(row_of(k), wst, ubins, hacc + grp * hstride + lane, shift, mask);
                    } else {
                        // accumulators in the unit's output row (global, L2): read, update through the scratch, write back
                        double * const sc0 = hacc + lane, * const sc1 = hacc + hstride + lane;
                        for (; k < n_int; k += 64) {
                            const bool two = k + 32 < n_int;                  // lanes of one warp agree except in the last group
                            for (int b = 0; b < bins; ++b) {
                                sc0[b * 32] = out[hist0 + k * bins + b];
                                if (two) sc1[b * 32] = out[hist0 + (k + 32) * bins + b];
                            }
                            if (two) hist_round2(row_of(k), row_of(k + 32), wst, ubins, sc0, sc1, shift, mask);
                            else hist_round(row_of(k), wst, ubins, sc0, shift, mask);
                            for (int b = 0; b < bins; ++b) {
                                out[hist0 + k * bins + b] = sc0[b * 32];
                                if (two) out[hist0 + (k + 32) * bins + b] = sc1[b * 32];
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }

        // the unit's partial row: base columns by the xor tree over lanes (the warp stage of block_reduce) ...
        double v[kBaseCols];
        v[col::max_lw] = max_lw;
        v[col::s0] = s0;
        v[col::s00] = s00;
        v[col::n_neginf] = static_cast<double>(n_neginf);
        const bool any_int = imin <= imax;
        v[col::neg_imin] = any_int ? -static_cast<double>(imin) : dm::neg_inf();
        v[col::imax] = any_int ? static_cast<double>(imax) : dm::neg_inf();
        v[col::int_oor] = (any_int && (imin < lo || static_cast<long long>(imax) >= static_cast<long long>(lo) + bins)) ? 1.0 : 0.0;
        v[col::n_nan] = static_cast<double>(n_nan);
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < kBaseCols; ++j) {
            double x = v[j];
            const bool is_max = (kMaxColsMask >> j) & 1ull;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double y = __shfl_xor_sync(0xffffffffu, x, off);
                x = is_max ? fmax(x, y) : x + y;
            }
            if (static_cast<int>(lane) == j) mine = x;
        }
        if (static_cast<int>(lane) < kBaseCols) out[lane] = mine;
        // ... and the row sums straight from the accumulators (already complete per row: no reduction over lanes)
        for (int k = static_cast<int>(lane); k < 2 * n_real; k += 32) out[kBaseCols + k] = macc[k];
        if (!hacc_global) {
            for (int k = static_cast<int>(lane); k < n_int * bins; k += 32) {
                const int row = k / bins, bin = k % bins;
                out[hist0 + k] = hacc[(row / 32) * hstride + bin * 32 + (row % 32)];
            }
        }
        __syncwarp();
    }
}

}  // namespace engine
}  // namespace cpprob
#endif  // CPPROB_B200_STAGED_KERNELS_CUH
