// cpprob-b200: model-independent posterior reduction kernels over the SoA trace rows, and the
// chunk-partial merge.  Together with the sums formed inside k_sis_fused / k_sis_rows they replace
// the reference's CPU estimator pass (/root/reference: include/cpprob/postprocess/
// stats_printer.hpp:88-120 keyed by (address id, k-th occurrence); empirical_distribution.hpp:52-81
// mean / variance, :30-40 distribution, :117-143 max-shifted logsumexp).
#ifndef CPPROB_B200_REDUCE_KERNELS_CUH
#define CPPROB_B200_REDUCE_KERNELS_CUH

#include "sis_kernels.cuh"
#include "staged_kernels.cuh"

namespace cpprob {
namespace engine {

__global__ void __launch_bounds__(kBlock) k_init_int_extra(int_extra * __restrict__ x, unsigned n)
{
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) x[i] = int_extra{0x7fffffff, static_cast<int>(0x80000000u)};
}

// ------------------------------------------------------------------------------------------------
// K4a k_rows_moments: S1 = sum w x, S2 = sum w x^2 for 32 real rows of one sub-chunk, in the canonical order of
// staged_kernels.cuh.  1-D grid of n_sub_chunks * ceil(n_real / 32) CTAs, row group fastest.  Warp s of the CTA is warp
// slot s: for each of its 16 rounds it loads the round's 32 particles of the group's 32 rows (32 coalesced 256-byte
// loads in flight per warp) into its staging area, and lane q adds the round to the sums of row q with the same
// moments_round the staged kernel uses.  The 8 slots are then added in slot order.  Each row element is read from
// HBM once; w comes from L2 after its first read.
// Output: partials[sub_chunk][kBaseCols + 2*row + {0,1}].
// ------------------------------------------------------------------------------------------------
constexpr unsigned kRowsMomentsSmem = kWarps * (32u * kStageRealStride + 32u) * sizeof(double) + kWarps * 64u * sizeof(double);

__global__ void __launch_bounds__(kBlock) k_rows_moments(const double * __restrict__ real_rows, const double * __restrict__ w,
                                                         unsigned long long stride, unsigned long long n_particles,
                                                         int n_real, double * __restrict__ partials, int n_cols)
{
    extern __shared__ double cpprob_zig_shared[];
    const unsigned lane = threadIdx.x & 31u, slot = threadIdx.x >> 5;
    double * const stage = cpprob_zig_shared + slot * (32u * kStageRealStride + 32u);
    double * const wst = stage + 32u * kStageRealStride;
    double * const slot_sums = cpprob_zig_shared + kWarps * (32u * kStageRealStride + 32u);      // [kWarps][32][2]
    const unsigned n_groups = static_cast<unsigned>((n_real + 31) / 32);
    const unsigned c = blockIdx.x / n_groups;
    const int row0 = static_cast<int>(blockIdx.x % n_groups) * 32;
    const unsigned long long base = static_cast<unsigned long long>(c) * kSubChunk;
    const unsigned long long left = n_particles - base;
    const unsigned n_here = left < kSubChunk ? static_cast<unsigned>(left) : kSubChunk;
    const int rows_here = min(32, n_real - row0);

    double s1 = 0.0, s2 = 0.0;
    for (unsigned round = 0; round < kSubChunk / kBlock; ++round) {
        const unsigned i0 = round * kBlock + slot * 32u;              // first particle of this slot's round within the sub-chunk
        if (i0 >= n_here) break;
        const bool valid = i0 + lane < n_here;
        const unsigned long long colidx = base + i0 + lane;
        wst[lane] = valid ? w[colidx] : 0.0;
        const double * __restrict__ src = real_rows + static_cast<unsigned long long>(row0) * stride + colidx;
        if (valid && rows_here == 32) {
#pragma unroll 16
            for (int q = 0; q < 32; ++q) stage[q * kStageRealStride + lane] = __ldcs(src + static_cast<unsigned long long>(q) * stride);
        } else {
            for (int q = 0; q < 32; ++q) {
                stage[q * kStageRealStride + lane] = (valid && q < rows_here) ? __ldcs(src + static_cast<unsigned long long>(q) * stride) : 0.0;
            }
        }
        __syncwarp();
        moments_round(stage + lane * kStageRealStride, wst, s1, s2);
        __syncwarp();
    }
    slot_sums[(slot * 32u + lane) * 2u] = s1;
    slot_sums[(slot * 32u + lane) * 2u + 1u] = s2;
    __syncthreads();
    if (threadIdx.x < 64u && static_cast<int>(threadIdx.x >> 1) < rows_here) {
        double r = slot_sums[threadIdx.x];
#pragma unroll
        for (unsigned sl = 1; sl < kWarps; ++sl) r = __dadd_rn(r, slot_sums[sl * 64u + threadIdx.x]);     // slots in slot order
        partials[static_cast<size_t>(c) * n_cols + kBaseCols + 2 * row0 + threadIdx.x] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// K4b k_rows_hist: weighted histogram sum_i w_i [x_i == lo + b], b < bins_here <= kRowsHistBins, for 32 int rows of one
// sub-chunk, canonical order as above (hist_round of staged_kernels.cuh).  1-D grid of n_sub_chunks * ceil(n_int / 32)
// CTAs.  Wider windows are covered by several launches with shifted `lo` (bin_offset selects the output columns).
// States are staged as bytes relative to `lo`; anything outside [0, 255) is staged as 255 and matches no bin.
// Output: partials[sub_chunk][hist_col0 + row*hist_bins + bin_offset + b].
// ------------------------------------------------------------------------------------------------
constexpr int kRowsHistBins = 16;

__global__ void __launch_bounds__(kBlock) k_rows_hist(const int * __restrict__ int_rows, const double * __restrict__ w,
                                                      unsigned long long stride, unsigned long long n_particles, int n_int,
                                                      long long lo, int bin_offset, int bins_here, int hist_bins, int hist_col0,
                                                      double * __restrict__ partials, int n_cols)
{
    __shared__ __align__(16) unsigned char stage_all[kWarps][32 * kStageIntStride];
    __shared__ __align__(16) double wst_all[kWarps][32];
    __shared__ double acc_all[kWarps][kRowsHistBins + 1][32];          // bin-major: see hist_acc_stride
    const unsigned lane = threadIdx.x & 31u, slot = threadIdx.x >> 5;
    unsigned char * const stage = stage_all[slot];
    double * const wst = wst_all[slot];
    double * const acc = &acc_all[slot][0][lane];
    const unsigned n_groups = static_cast<unsigned>((n_int + 31) / 32);
    const unsigned c = blockIdx.x / n_groups;
    const int row0 = static_cast<int>(blockIdx.x % n_groups) * 32;
    const unsigned long long base = static_cast<unsigned long long>(c) * kSubChunk;
    const unsigned long long left = n_particles - base;
    const unsigned n_here = left < kSubChunk ? static_cast<unsigned>(left) : kSubChunk;
    const int rows_here = min(32, n_int - row0);
    const unsigned lo32 = static_cast<unsigned>(static_cast<int>(lo));

    for (int b = 0; b <= bins_here; ++b) acc[b * 32] = 0.0;
    for (unsigned round = 0; round < kSubChunk / kBlock; ++round) {
        const unsigned i0 = round * kBlock + slot * 32u;
        if (i0 >= n_here) break;
        const bool valid = i0 + lane < n_here;
        const unsigned long long colidx = base + i0 + lane;
        wst[lane] = valid ? w[colidx] : 0.0;
        const int * __restrict__ src = int_rows + static_cast<unsigned long long>(row0) * stride + colidx;
        if (valid && rows_here == 32) {
#pragma unroll 16
            for (int q = 0; q < 32; ++q) {
                const unsigned x = static_cast<unsigned>(__ldcs(src + static_cast<unsigned long long>(q) * stride)) - lo32;
                stage[q * kStageIntStride + lane] = static_cast<unsigned char>(x < 255u ? x : 255u);
            }
        } else {
            for (int q = 0; q < 32; ++q) {
                unsigned x = 255u;
                if (valid && q < rows_here) x = static_cast<unsigned>(__ldcs(src + static_cast<unsigned long long>(q) * stride)) - lo32;
                stage[q * kStageIntStride + lane] = static_cast<unsigned char>(x < 255u ? x : 255u);
            }
        }
        __syncwarp();
        hist_round(stage + lane * kStageIntStride, wst, static_cast<unsigned>(bins_here), acc);   // `lo` already is the first bin of this launch
        __syncwarp();
    }
    __syncthreads();
    for (unsigned t = threadIdx.x; t < 32u * static_cast<unsigned>(bins_here); t += kBlock) {
        const unsigned q = t / static_cast<unsigned>(bins_here), b = t % static_cast<unsigned>(bins_here);
        if (static_cast<int>(q) < rows_here && bin_offset + static_cast<int>(b) < hist_bins) {
            double r = acc_all[0][b][q];
#pragma unroll
            for (unsigned sl = 1; sl < kWarps; ++sl) r = __dadd_rn(r, acc_all[sl][b][q]);       // slots in slot order
            partials[static_cast<size_t>(c) * n_cols + hist_col0 + (row0 + static_cast<int>(q)) * hist_bins + bin_offset + static_cast<int>(b)] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_fold_rows: super-chunk rows.  out[s][c] = rows [s*per, min((s+1)*per, n_rows)) of `in` combined in row
// order by one thread (max columns with fmax, the others with +).  What a rank hands to the gather.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_fold_rows(const double * __restrict__ in, unsigned n_rows, unsigned per, int n_cols,
                                                      unsigned long long max_mask, double * __restrict__ out, unsigned n_out_rows)
{
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x;
    if (i >= static_cast<unsigned long long>(n_out_rows) * n_cols) return;
    const unsigned s = static_cast<unsigned>(i / n_cols);
    const int c = static_cast<int>(i % n_cols);
    const bool is_max = c < 64 && ((max_mask >> c) & 1ull);
    const unsigned long long r0 = static_cast<unsigned long long>(s) * per;
    const unsigned long long r1 = r0 + per < n_rows ? r0 + per : n_rows;
    double v = is_max ? dm::neg_inf() : 0.0;
    for (unsigned long long r = r0; r < r1; ++r) {
        const double x = in[r * n_cols + c];
        v = is_max ? fmax(v, x) : v + x;
    }
    out[i] = v;
}

// ------------------------------------------------------------------------------------------------
// k_merge_columns: out[col] = reduce over chunks (rows of `partials`) in a fixed order.
// grid = n_cols CTAs.  Thread t folds rows t, t+256, ... sequentially, then the fixed CTA tree.
// The input is the concatenation of every rank's partials in chunk order, so the result is
// bit-identical on every rank and for every GPU count.
// ------------------------------------------------------------------------------------------------
// out[n_cols] = *m_ref_dev (when given): the run's reference log-weight comes back with the sums, in one copy.
__global__ void __launch_bounds__(kBlock) k_merge_columns(const double * __restrict__ partials, unsigned n_rows, int n_cols,
                                                          unsigned long long max_mask, double * __restrict__ out,
                                                          const double * __restrict__ m_ref_dev)
{
    if (blockIdx.x == 0 && threadIdx.x == 32 && m_ref_dev) out[n_cols] = *m_ref_dev;
    __shared__ double smem[kWarps];
    const int c = blockIdx.x;
    const bool is_max = c < 64 && ((max_mask >> c) & 1ull);
    double v[1];
    v[0] = is_max ? dm::neg_inf() : 0.0;
    for (unsigned r = threadIdx.x; r < n_rows; r += kBlock) {
        const double x = partials[static_cast<size_t>(r) * n_cols + c];
        v[0] = is_max ? fmax(v[0], x) : v[0] + x;
    }
    const double res = block_reduce<1>(v, is_max ? 1ull : 0ull, smem);
    if (threadIdx.x == 0) out[c] = res;
}

// The same merge straight from the output of an all-gather: `world` segments of rows_per_rank rows each, segment r
// holding the first[r+1] - first[r] rows of rank r (the rest of a segment is padding that is never read).  Logical
// row i lives at segment r = the rank that owns it, position i - first[r]; every thread walks its rows in the same
// order as k_merge_columns, so the result has the same bits — no compaction copy in between.
constexpr int kMaxMergeRanks = 64;
struct gather_layout {
    unsigned first[kMaxMergeRanks + 1];      // first[r] = index of rank r's first logical row; first[world] = total rows
    unsigned world;
    unsigned rows_per_rank;
};

// ------------------------------------------------------------------------------------------------
// The exchange of a multi-GPU inference over peer memory (NVLink / NVSwitch), fused with the kernels on either side of it.
// Every rank owns a window: [kPeerFlagBytes of flags][two gather buffers]; every rank holds a pointer to every other rank's
// window (cudaIpcOpenMemHandle between processes, cudaDeviceEnablePeerAccess within one).
//   k_push_rows      (the producer's last step): this rank's partial rows are stored straight into its segment of EVERY
//                    rank's gather buffer (its own included), then — by the last CTA to finish, after a system-wide fence
//                    — the inference's epoch into its slot of every rank's flag array;
//   k_merge_columns_gathered (the consumer): waits until all `world` flags of its own window carry the epoch, then merges
//                    the gathered rows where they lie.
// No NCCL kernel, no host round trip: the exchange costs one small kernel and the NVLink latency of a store + a flag.
// Two gather buffers alternate by the epoch's parity: a rank can run at most one inference ahead of the slowest peer (its
// next merge waits for that peer's push, which that peer's stream orders behind its own previous merge).
// ------------------------------------------------------------------------------------------------
constexpr unsigned kPeerFlagBytes = 1024;     // 64 epoch slots (unsigned long long) + an error word + the push kernel's CTA counter
constexpr unsigned kPeerErrorWord = 64;       // index (in unsigned long long) of the error word
constexpr unsigned kPeerDoneWord = 65;        // index of the CTA-done counter of k_push_rows (local use only)

struct peer_targets {
    unsigned char * window[kMaxMergeRanks];   // every rank's window as seen from this device
    unsigned world;
    unsigned rank;
};

// what a producing kernel needs to push: the windows, where this inference's gather buffer starts in them, the length
// of a rank's segment, and the epoch to publish (world == 0: nothing to push)
struct peer_push {
    peer_targets t;
    unsigned long long buffer_offset_bytes;
    unsigned long long segment_doubles;
    unsigned long long epoch;
};

__device__ __forceinline__ double * peer_segment(const peer_push & pp, unsigned p)
{
    return reinterpret_cast<double *>(pp.t.window[p] + pp.buffer_offset_bytes) + static_cast<unsigned long long>(pp.t.rank) * pp.segment_doubles;
}

// Called by every thread of every CTA after its stores: the last CTA to get here publishes the epoch in every rank's
// window (the pattern of the threadFenceReduction sample, at system scope).
__device__ __forceinline__ void peer_publish(const peer_push & pp)
{
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    unsigned long long * const mine = reinterpret_cast<unsigned long long *>(pp.t.window[pp.t.rank]);
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(mine + kPeerDoneWord, 1ull);
        last = done + 1 == gridDim.x;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) mine[kPeerDoneWord] = 0ull;             // ready for the next launch
    __threadfence_system();
    if (threadIdx.x < pp.t.world) {
        volatile unsigned long long * flag = reinterpret_cast<volatile unsigned long long *>(pp.t.window[threadIdx.x]) + pp.t.rank;
        *flag = pp.epoch;
    }
}

// rows that already lie in a buffer of this rank (row path): copy + publish
__global__ void __launch_bounds__(kBlock) k_push_rows(const double * __restrict__ rows, unsigned long long n_doubles, const __grid_constant__ peer_push pp)
{
    const unsigned long long i0 = blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x;
    const unsigned long long step = static_cast<unsigned long long>(gridDim.x) * kBlock;
    for (unsigned long long i = i0; i < n_doubles; i += step) {
        const double v = rows[i];
        for (unsigned p = 0; p < pp.t.world; ++p) peer_segment(pp, p)[i] = v;
    }
    peer_publish(pp);
}

// ------------------------------------------------------------------------------------------------
// k_fold_units: the particle kernels' per-unit rows -> the rows a rank hands on, in one step, and (multi-GPU, peer
// exchange) straight into every rank's gather buffer.  Output row s, column c = kernel rows [s*per_super, (s+1)*per_super)
// combined in row order, each kernel row being its `per_unit` unit rows of `units` combined in row order — the very
// additions a fold of the unit rows followed by k_fold_rows would make (same bits, round 2's first version did exactly
// that), without the intermediate array and the second launch.
// A group of g = 2^k lanes (g >= per_super, at most 32; the host picks it) forms one output element: lane j of the group
// loads and folds the unit rows of kernel row r0 + j (+ g, + 2g ... for super-chunks of more than 32 rows) — the loads of
// a long super-chunk (64 rows x 8 units when 8e9 particles are spread over 8 GPUs) run side by side instead of one after
// the other in a single thread — and the rows' values are then added IN ROW ORDER through shuffles, every lane of the
// group forming the same sum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_fold_units(const double * __restrict__ units, unsigned n_rows, int per_unit, int nv, unsigned per_super,
                                                       int n_cols, unsigned long long max_mask, double * __restrict__ out, unsigned n_out_rows,
                                                       unsigned group, const __grid_constant__ peer_push pp)
{
    const unsigned long long n_out = static_cast<unsigned long long>(n_out_rows) * n_cols;
    const unsigned long long slot = (blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x) / group;   // output element of this group
    const unsigned sub = threadIdx.x % group;
    // (whole warps stay together: an element beyond the end just idles through the shuffles)
    const bool live = slot < n_out;
    const unsigned long long i = live ? slot : 0ull;
    const unsigned s = static_cast<unsigned>(i / n_cols);
    const int c = static_cast<int>(i % n_cols);
    const bool is_max = c < 64 && ((max_mask >> c) & 1ull);
    const unsigned long long r0 = static_cast<unsigned long long>(s) * per_super;
    const unsigned long long r1 = r0 + per_super < n_rows ? r0 + per_super : n_rows;
    double acc = is_max ? dm::neg_inf() : 0.0;                      // (k_fold_rows starts from the identity too: same bits down to -0.0)
    const unsigned trips = (per_super + group - 1) / group;         // the same for every lane of the grid: the shuffles below are warp-wide
    for (unsigned k = 0; k < trips; ++k) {
        const unsigned long long base = r0 + static_cast<unsigned long long>(k) * group;
        const unsigned long long r = base + sub;
        double x = 0.0;
        if (live && r < r1) {
            const double * p = units + (r * per_unit) * nv + c;
            x = p[0];
#pragma unroll 8
            for (int w = 1; w < per_unit; ++w) {
                const double y = p[static_cast<size_t>(w) * nv];
                x = is_max ? fmax(x, y) : x + y;
            }
        }
        const unsigned here = base >= r1 ? 0u : (r1 - base < group ? static_cast<unsigned>(r1 - base) : group);
        for (unsigned j = 0; j < group; ++j) {                      // (uniform trip count across the warp; `here` masks the tail)
            const double v = __shfl_sync(0xffffffffu, x, static_cast<int>(j), static_cast<int>(group));
            if (j < here) acc = per_super == 1u ? v : (is_max ? fmax(acc, v) : acc + v);   // (one row per output row: the row itself)
        }
    }
    if (live && sub == 0u) {
        out[i] = acc;
        for (unsigned p = 0; p < pp.t.world; ++p) peer_segment(pp, p)[i] = acc;
    }
    if (pp.t.world) peer_publish(pp);
}

// the consumer's wait: all `world` slots of this rank's own flag array reach `epoch` (or ~20 s pass: a peer died; the
// error word is raised, the merge goes on over whatever is there and the host reports the failure)
__device__ __forceinline__ void peer_wait(const unsigned long long * flags_in, unsigned world, unsigned long long epoch)
{
    if (threadIdx.x < world) {
        const volatile unsigned long long * f = flags_in + threadIdx.x;
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while (*f < epoch) {
            if ((++spins & 0x3ffu) == 0u) {
                unsigned long long t1 = 0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {
                    const_cast<unsigned long long *>(flags_in)[kPeerErrorWord] = epoch;
                    break;
                }
            }
        }
    }
    __threadfence_system();
    __syncthreads();
}

// With `peer_flags` the rows were pushed by the peers themselves (k_push_rows): wait for their epoch flags first and read
// the rows past L1 (they were written by other GPUs).
__global__ void __launch_bounds__(kBlock) k_merge_columns_gathered(const double * gathered, const __grid_constant__ gather_layout lay,
                                                                   int n_cols, unsigned long long max_mask, double * __restrict__ out,
                                                                   const double * __restrict__ m_ref_dev,
                                                                   const unsigned long long * peer_flags, unsigned long long epoch)
{
    if (peer_flags) peer_wait(peer_flags, lay.world, epoch);
    if (blockIdx.x == 0 && threadIdx.x == 32 && m_ref_dev) out[n_cols] = *m_ref_dev;
    if (blockIdx.x == 0 && threadIdx.x == 33 && peer_flags) out[n_cols + 1] = peer_flags[kPeerErrorWord] == epoch ? 1.0 : 0.0;
    __shared__ double smem[kWarps];
    const int c = blockIdx.x;
    const bool is_max = c < 64 && ((max_mask >> c) & 1ull);
    const unsigned n_rows = lay.first[lay.world];
    double v[1];
    v[0] = is_max ? dm::neg_inf() : 0.0;
    unsigned rank = 0;
    for (unsigned r = threadIdx.x; r < n_rows; r += kBlock) {
        while (r >= lay.first[rank + 1]) ++rank;                  // rows only grow: the owner is found by walking on
        const size_t phys = static_cast<size_t>(rank) * lay.rows_per_rank + (r - lay.first[rank]);
        const double x = __ldcg(gathered + phys * n_cols + c);
        v[0] = is_max ? fmax(v[0], x) : v[0] + x;
    }
    const double res = block_reduce<1>(v, is_max ? 1ull : 0ull, smem);
    if (threadIdx.x == 0) out[c] = res;
}

// ------------------------------------------------------------------------------------------------
// K3 k_row_base: weights w = exp(log_w - m_ref) and the base sums of one sub-chunk from its log_w
// column (16 B of HBM traffic per particle).  Used after k_sis_rows and for externally supplied
// records (replay / StatsPrinter-on-device).  `extras` (may be null) carries the int bookkeeping that
// k_sis_rows accumulated for the sub-chunk (min / max of the stored ints; the out-of-window flag derives from them).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_row_base(const double * __restrict__ logw, unsigned long long n_particles, unsigned chunk,
                                                     const double * __restrict__ m_ref_ptr, double * __restrict__ w_out,
                                                     const int_extra * __restrict__ extras, long long hist_lo, int hist_bins,
                                                     double * __restrict__ partials, int n_cols)
{
    __shared__ double smem[kWarps * kBaseCols];
    const unsigned c = blockIdx.x;
    const double m_ref = *m_ref_ptr;
    const unsigned long long base = static_cast<unsigned long long>(c) * chunk;
    const unsigned long long left = n_particles - base;
    const unsigned n_here = left < chunk ? static_cast<unsigned>(left) : chunk;
    double v[kBaseCols];
#pragma unroll
    for (int j = 0; j < kBaseCols; ++j) v[j] = 0.0;
    v[col::max_lw] = dm::neg_inf();
    v[col::neg_imin] = dm::neg_inf();
    v[col::imax] = dm::neg_inf();
    for (unsigned i = threadIdx.x; i < n_here; i += kBlock) {
        const double lw = logw[base + i];
        const double wi = dm::exp_weight(lw - m_ref);
        w_out[base + i] = wi;
        v[col::max_lw] = lw > v[col::max_lw] ? lw : v[col::max_lw];
        v[col::s0] += wi;
        v[col::s00] = fma(wi, wi, v[col::s00]);
        v[col::n_neginf] += is_neg_inf(lw) ? 1.0 : 0.0;
        v[col::n_nan] += is_nan(lw) ? 1.0 : 0.0;
    }
    double r = block_reduce<kBaseCols>(v, kMaxColsMask, smem);
    if (extras != nullptr) {
        const int_extra x = extras[c];
        if (threadIdx.x == col::neg_imin) r = x.vmin <= x.vmax ? -static_cast<double>(x.vmin) : dm::neg_inf();
        if (threadIdx.x == col::imax) r = x.vmin <= x.vmax ? static_cast<double>(x.vmax) : dm::neg_inf();
        // non-zero when some int predict of the sub-chunk fell outside the histogram window [lo, lo + bins)
        if (threadIdx.x == col::int_oor) r = (x.vmin <= x.vmax && (x.vmin < hist_lo || x.vmax >= hist_lo + hist_bins)) ? 1.0 : 0.0;
    }
    if (threadIdx.x < kBaseCols) partials[static_cast<size_t>(c) * n_cols + threadIdx.x] = r;
}

// k_max_array: max of an array (for choosing m_ref from external log-weights).  One CTA.
__global__ void __launch_bounds__(kBlock) k_max_array(const double * __restrict__ x, unsigned long long n, double * __restrict__ out)
{
    __shared__ double smem[kWarps];
    double v[1] = {dm::neg_inf()};
    for (unsigned long long i = threadIdx.x; i < n; i += kBlock) v[0] = fmax(v[0], x[i]);
    const double r = block_reduce<1>(v, 1ull, smem);
    if (threadIdx.x == 0) out[0] = (r > -1.0e300 && r < 1.0e300) ? r : 0.0;
}

// ------------------------------------------------------------------------------------------------
// K0 micro-benchmarks: the roofline denominators (MEASURED_PEAKS.json has no FP64 entry).
// ------------------------------------------------------------------------------------------------
// Register-only DFMA chains: 8 independent accumulators per thread, `iters` x 8 x 4 DFMA each.
__global__ void __launch_bounds__(kBlock) k_dfma_peak(double * __restrict__ out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * kBlock + threadIdx.x] = s;   // never true; keeps the chain alive
}

// Latency probe: C independent DFMA chains per thread (C = 1, 2, 4, 8), `iters` x 32 DFMA per chain.
template<int C>
__global__ void __launch_bounds__(kBlock) k_dfma_chains(double * __restrict__ out, int iters, double a, double b)
{
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = fma(x[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    if (s == 123.456) out[blockIdx.x * kBlock + threadIdx.x] = s;
}

// Issue-model probe: the same 8 DFMA chains with NI independent integer (ALU-pipe) instructions per DFMA
// interleaved.  If time does not grow with NI <= 1 the FP64 pipe co-issues with the ALU pipe; if it
// grows by ~50% per NI, an FP64 warp instruction holds the issue port for both of its cycles.
template<int NI>
__global__ void __launch_bounds__(kBlock) k_issue_probe(double * __restrict__ out, int iters, double a, double b, unsigned m)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                i0 = (i0 ^ m) + i1; i1 = (i1 ^ m) + i2; i2 = (i2 ^ m) + i3; i3 = (i3 ^ m) + i4;
                i4 = (i4 ^ m) + i5; i5 = (i5 ^ m) + i6; i6 = (i6 ^ m) + i7; i7 = (i7 ^ m) + i0;
            }
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    const unsigned t = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if (s == 123.456 || t == 0x12345u) out[blockIdx.x * kBlock + threadIdx.x] = s + t;
}

// Streaming store of doubles (HBM write roofline for the trace rows).
__global__ void __launch_bounds__(kBlock) k_store_peak(double2 * __restrict__ dst, unsigned long long n2, double v)
{
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kBlock) + threadIdx.x; i < n2;
         i += static_cast<unsigned long long>(gridDim.x) * kBlock) {
        dst[i] = make_double2(v, v);
    }
}

}  // namespace engine
}  // namespace cpprob
#endif  // CPPROB_B200_REDUCE_KERNELS_CUH
