/* cpprob-b200: C ABI of the B200 sequential-importance-sampling engine (libcpprob_sis.so).
 *
 * CPProb has no FFI of its own: its boundary for this path is the header-only C++ template API
 * (/root/reference: include/cpprob/cpprob.hpp:173-203 `cpprob::inference`, :68-106 sample/observe/
 * predict; include/cpprob/postprocess/stats_printer.hpp:22-121).  The C++14 mirror of that API in
 * include/cpprob/cpprob.hpp is a thin header layer over the entry points declared here; each entry
 * point names the reference code whose work it takes over.  Plain pointers and sizes only, no
 * exceptions cross this boundary: every call returns 0 on success or a negative CPPROB_SIS_E* code,
 * and cpprob_sis_last_error() returns the message for the calling thread.
 *
 * There is no CPU fallback: if no CUDA device is usable cpprob_sis_create fails.
 */
#ifndef CPPROB_SIS_H
#define CPPROB_SIS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPPROB_SIS_ABI_VERSION 3

enum {
    CPPROB_SIS_OK = 0,
    CPPROB_SIS_EINVAL = -1,    /* bad argument */
    CPPROB_SIS_ECUDA = -2,     /* CUDA runtime error (message has the details) */
    CPPROB_SIS_ENOMODEL = -3,  /* unknown model id / name */
    CPPROB_SIS_EIO = -4,       /* posterior file could not be written */
    CPPROB_SIS_ERANGE = -5,    /* int predict window too wide for the on-device histogram */
    CPPROB_SIS_ENOMEM = -6,
    CPPROB_SIS_ENCCL = -7      /* NCCL missing or an NCCL call failed (multi-GPU entry points only) */
};

typedef struct cpprob_sis_engine cpprob_sis_engine;

typedef struct cpprob_sis_config {
    int device;              /* CUDA device ordinal */
    uint64_t seed;           /* Philox key; the reference seeds mt19937 from random_device
                                (src/cpprob/utils.cpp:16-20), so it has no counterpart */
    uint64_t max_batch;      /* particles per trace batch when rows are materialised (0 = default) */
    int blocks_per_sm;       /* persistent grid = blocks_per_sm x SM count (0 = default) */
} cpprob_sis_config;

/* ---- engine lifetime ------------------------------------------------------------------------ */
int cpprob_sis_abi_version(void);
const char * cpprob_sis_last_error(void);
int cpprob_sis_create(const cpprob_sis_config * cfg, cpprob_sis_engine ** out);
void cpprob_sis_destroy(cpprob_sis_engine * e);
/* A new Philox key for the runs that follow.  An engine keeps its streams, buffers and tables between calls: a host API
 * that serves many inference() calls (include/cpprob/cpprob.hpp) keeps ONE engine per GPU and re-seeds it, instead of
 * paying the context / allocation / table-upload cost on every call. */
int cpprob_sis_set_seed(cpprob_sis_engine * e, uint64_t seed);

/* ---- model registry --------------------------------------------------------------------------
 * A model is a device functor (include/models/models.hpp) compiled into a set of kernel
 * instantiations; the registry pairs a name with their launchers.  Built-in: the models of
 * /root/reference include/models/models.hpp:22-35,67-80,114-141 and src/models/gaussian.cpp:6-17.
 * Plugins (user .cu files using CPPROB_SIS_REGISTER_MODEL) add theirs at load time. */
struct cpprob_sis_model_vtable;
int cpprob_sis_register_model(const struct cpprob_sis_model_vtable * vt);
int cpprob_sis_model_count(void);
const char * cpprob_sis_model_name(int model_id);
int cpprob_sis_find_model(const char * name);     /* id >= 0, or CPPROB_SIS_ENOMODEL */

/* ---- trace structure -------------------------------------------------------------------------
 * Takes over TraceInfer::register_addr_predict (include/cpprob/trace.hpp:37-41) and the
 * int/real routing of StateInfer::add_predict (include/cpprob/state.hpp:312-326): one host-side
 * probe execution records, in program order, the address id and kind of every predict statement. */
typedef struct cpprob_sis_slot {
    int is_int;      /* 1: value goes to <out>.int, 0: to <out>.real */
    int id;          /* address id = line number in <out>.ids */
    int k;           /* occurrence index of this id within one trace (StatsPrinter's key) */
    int row;         /* (first) row inside the int / real SoA block */
    int width;       /* 1 for a scalar predict; N for a vector predict = N consecutive real rows, `(id [v0 .. vN-1])` */
} cpprob_sis_slot;

typedef struct cpprob_sis_structure {
    int n_ids;
    int n_slots;
    int n_real;
    int n_int;
    int n_samples;
    const char * const * ids;          /* [n_ids], engine-owned, valid until the next describe/run */
    const cpprob_sis_slot * slots;     /* [n_slots], program order */
} cpprob_sis_structure;

int cpprob_sis_describe(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                        cpprob_sis_structure * out);

/* ---- inference --------------------------------------------------------------------------------
 * Posterior estimators.  Everything StatsPrinter prints (mean / variance per real (id,k);
 * distribution / MAP / num points per int (id,k)) plus max log-weight, log-sum-exp, log-evidence
 * and ESS, which the reference does not compute (north_star additions).  Pointers are engine-owned
 * and stay valid until the next call on the same engine. */
typedef struct cpprob_sis_stats {
    uint64_t n_particles;
    uint64_t n_neg_inf;          /* particles with log_w == -inf */
    uint64_t n_nan;
    double m_ref;                /* reference log-weight the sums are relative to */
    double max_log_w;
    double log_sum_exp;          /* log sum_i exp(log_w_i)  (empirical_distribution.hpp:117-143) */
    double log_evidence;         /* log_sum_exp - log n */
    double ess;                  /* (sum w)^2 / sum w^2 */
    int n_real;
    int n_int;
    const double * real_mean;    /* [n_real]  sum w~ x            (empirical_distribution.hpp:52-71) */
    const double * real_var;     /* [n_real]  sum w~ x^2 - mean^2 (:78-81) */
    long long int_lo;            /* first histogram bin value */
    int int_bins;
    const double * int_prob;     /* [n_int][int_bins]  sum w~ [x == int_lo + b]   (:30-40) */
    const long long * int_map;   /* [n_int] argmax (first maximum, as std::max_element :47-50) */
    int n_cols;
    const double * sums;         /* [n_cols] merged raw sums (kBaseCols, then S1,S2 per real row, then bins) */
    double device_ms;            /* CUDA-event time of the particle + reduction kernels of this call */
    uint64_t kernel_launches;    /* kernels launched by this call */
    int passes;                  /* 1, or 2 if m_ref had to be re-based */
    int path;                    /* CPPROB_SIS_PATH_*: where the traces lived while the estimators were formed */
    double particle_ms;          /* the part of device_ms that is this GPU's own particle pass (pilot, particle kernels, row
                                    reductions): without the merge, which on several GPUs includes the wait for the peers */
} cpprob_sis_stats;

enum {
    CPPROB_SIS_PATH_FUSED = 0,   /* registers (k_sis_fused): <= 4 real predicts, no int predicts, nothing emitted */
    CPPROB_SIS_PATH_STAGED = 1,  /* per-warp shared-memory staging areas (k_sis_staged): longer traces, nothing emitted */
    CPPROB_SIS_PATH_ROWS = 2     /* SoA rows in HBM (k_sis_rows + k_rows_moments / k_rows_hist): emitting runs, very long traces */
};

enum {
    CPPROB_SIS_EMIT_NONE = 0,    /* estimators only: no trace leaves the GPU */
    CPPROB_SIS_EMIT_ALL = 1      /* every particle's record is delivered / written */
};

typedef struct cpprob_sis_block {
    uint64_t first_particle;     /* global index of column 0 */
    uint64_t n;                  /* particles (columns) in this block */
    uint64_t stride;             /* elements between consecutive rows */
    int n_real, n_int;
    const double * real_rows;    /* [n_real][stride], pinned host memory */
    const int32_t * int_rows;    /* [n_int][stride] */
    const double * log_w;        /* [n] */
} cpprob_sis_block;

/* Called on the calling thread, once per trace block, in particle order.  Non-zero aborts the run. */
typedef int (*cpprob_sis_block_fn)(void * user, const cpprob_sis_block * blk);

typedef struct cpprob_sis_run_options {
    int emit;                    /* CPPROB_SIS_EMIT_* */
    int force_rows;              /* 1: use the row (SoA) path even when the fused kernel applies */
    cpprob_sis_block_fn on_block;/* receives the SoA trace blocks when emit == EMIT_ALL (may be NULL) */
    void * user;
} cpprob_sis_run_options;

/* Takes over the loop of cpprob::inference (cpprob.hpp:194-201) and the estimator pass of
 * StatsPrinter: runs n_particles weighted executions of the model on the engine's GPU. */
int cpprob_sis_run(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                   uint64_t n_particles, const cpprob_sis_run_options * opt, cpprob_sis_stats * out);

/* Same, and writes the reference's posterior files: one `([(id v) ...] logw)` line per particle to
 * <prefix>.real / <prefix>.int (StateInfer::finish_trace + dump_predicts, src/cpprob/state.cpp:
 * 193-202,262-267, grammar include/cpprob/serialization.hpp:41-46,71-98, scientific/precision 15),
 * <prefix>.ids (dump_ids, state.cpp:250-260) and removes the kinds that stayed empty
 * (finish_infer, state.cpp:164-180).  Files are appended to, as in the reference (ios::app).
 * Also writes <prefix>.stats with the on-device estimators.  With emit == CPPROB_SIS_EMIT_NONE the
 * per-particle records are not produced at all (no trace leaves the GPU): only <prefix>.ids and
 * <prefix>.stats are written. */
int cpprob_sis_infer_to_files(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                              uint64_t n_particles, const char * prefix, int emit /* CPPROB_SIS_EMIT_* */,
                              cpprob_sis_stats * out);

/* The same with the particles sharded over several GPUs of this process (engines[r] = rank r, one seed): every rank
 * formats its own particle range on its own GPU and writes it into its own byte range of <prefix>.real / .int (the
 * files are extended once, after a first pass that only measures each rank's text), so the files are byte for byte what
 * one GPU writes ("each rank streams its own particle range; the host writer concatenates in rank order").  .ids and
 * .stats are written once; the estimators are merged from the ranks' partial rows on engines[0]'s GPU. */
int cpprob_sis_infer_to_files_multi(cpprob_sis_engine * const * engines, int n_engines, int model_id, const double * obs, size_t n_obs,
                                    uint64_t n_particles, const char * prefix, cpprob_sis_stats * out);

/* Stage times of the last cpprob_sis_infer_to_files(..., EMIT_ALL) on this engine: the lines are formatted on
 * the GPU (`%.15e` exactly as the reference's ostream, cpprob_b200/csrc/text_format.cuh), only text crosses
 * PCIe, and the host appends it with concurrent pwrites.  kernel_ms: text kernels (CUDA events); copy_ms:
 * device->pinned-host copies (CUDA events); write_s: host wall time inside the file writes; bytes: text
 * written; fixups: lines the host had to re-format (undecidable last digit, probability ~1e-22 per value).
 * Any pointer may be NULL.  CPPROB_SIS_TEXT=host selects the older path (binary rows to the host, formatted
 * there by std::to_chars), which produces the same bytes. */
int cpprob_sis_text_stage_stats(cpprob_sis_engine * e, double * kernel_ms, double * copy_ms, double * write_s,
                                uint64_t * bytes, uint64_t * fixups);

/* ---- multi-GPU: shard / gather / merge ---------------------------------------------------------
 * Particles are i.i.d. (cpprob.hpp:194-201 has no inter-particle dependence), so rank r of `world`
 * takes the chunk range [r*C/world, (r+1)*C/world) of the C = ceil(n/32768) chunks and no data-path
 * collective is needed.  Each rank returns its per-chunk partial sums in DEVICE memory; the caller
 * all-gathers them in rank order (NCCL) and hands the concatenation to cpprob_sis_merge, which is
 * bit-identical on every rank and for every world size. */
/* The rows a rank hands to the gather (host arithmetic only).  Runs of up to 4096 chunks hand on their chunk
 * rows (times rows_per_chunk); larger runs group 2^k consecutive chunks into super-chunks (k from the run's size
 * only), a rank owns whole super-chunks and reduces each to one row before the gather: at most 4096 rows are
 * exchanged whatever the particle count, and the merged sums stay bit-identical for any world size.
 * rows_per_chunk: the value cpprob_sis_run_shard reported (1 or 8). */
int cpprob_sis_plan_rows(uint64_t n_particles_total, int rank, int world, int rows_per_chunk, uint32_t * row_first,
                         uint32_t * n_rows_local, uint32_t * n_rows_total);

typedef struct cpprob_sis_partials {
    double * device_ptr;         /* [n_chunks_local][n_cols] partial rows (see cpprob_sis_plan_rows), engine-owned */
    uint32_t n_chunks_local;     /* partial rows of this rank ... */
    uint32_t n_chunks_total;     /* ... of the whole run ... */
    uint32_t chunk_first;        /* ... and the index of this rank's first row */
    uint32_t rows_per_chunk;     /* rows the kernels wrote per 32768-particle chunk: 1 (fused kernel) or 8 (row path);
                                    argument of cpprob_sis_plan_rows */
    int n_cols;
    double m_ref;
    double device_ms;
    uint64_t kernel_launches;
} cpprob_sis_partials;

/* The shard of `rank`: pure host arithmetic, usable without a GPU. */
int cpprob_sis_plan_shard(uint64_t n_particles_total, int rank, int world, uint32_t * chunk_first,
                          uint32_t * n_chunks_local, uint32_t * n_chunks_total, uint64_t * first_particle,
                          uint64_t * n_local);

int cpprob_sis_run_shard(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                         uint64_t n_particles_total, int rank, int world,
                         const double * m_ref_override /* NULL: pilot */,
                         cpprob_sis_partials * out);

/* ---- multi-GPU inside the library: one exchange of the partial rows per inference (peer memory, or an NCCL all-gather) ----
 * north_star: "each rank reduces locally, and one small NCCL [collective] over NVLink combines the global max
 * log-weight, the sum of exp-weights and the weighted moment sums".  The collective is an all-gather of the per-
 * (super-)chunk partial rows (<= 4096 rows of n_cols doubles in all) on the engine's own stream, right behind the
 * kernels that wrote them; every rank then merges the gathered rows in place (k_merge_columns_gathered), which gives the
 * same bits on every rank and for every rank count.  The host synchronises once per inference.
 * NCCL is opened at run time (libnccl.so.2; CPPROB_SIS_NCCL_LIB overrides): nothing here is needed on one GPU.
 * How the rows travel: when every rank of the communicator can map every other rank's memory (one node, NVLink /
 * NVSwitch: CUDA IPC between processes, peer access within one), each rank's kernels store its rows straight into every
 * peer's gather buffer and raise a flag there, and the merge kernel waits for the flags — no collective kernel at all
 * (CPPROB_SIS_EXCHANGE_PEER).  Otherwise, or with the environment variable CPPROB_SIS_EXCHANGE=nccl, one ncclAllGather
 * (CPPROB_SIS_EXCHANGE_NCCL).  The communicator decides once, at cpprob_sis_comm_init, alike on every rank; the results
 * are the same bits either way.
 *
 * Process per GPU: rank 0 calls cpprob_sis_comm_get_id and hands the 128 bytes to the others by any means (MPI,
 * torch.distributed, a file); every rank calls cpprob_sis_comm_init on its engine (collective), then
 * cpprob_sis_run_dist with the same arguments (collective).  All engines must be created with ONE seed. */
#define CPPROB_SIS_COMM_ID_BYTES 128
int cpprob_sis_comm_get_id(void * id_out /* [CPPROB_SIS_COMM_ID_BYTES] */);
int cpprob_sis_comm_init(cpprob_sis_engine * e, const void * id, int rank, int world);
int cpprob_sis_comm_destroy(cpprob_sis_engine * e);
enum { CPPROB_SIS_EXCHANGE_NONE = 0, CPPROB_SIS_EXCHANGE_NCCL = 1, CPPROB_SIS_EXCHANGE_PEER = 2 };
/* which exchange the engine's communicator uses (NONE: rank 0 of 1) */
int cpprob_sis_comm_exchange(const cpprob_sis_engine * e);
int cpprob_sis_run_dist(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, uint64_t n_particles_total,
                        cpprob_sis_stats * out);
/* One process driving several GPUs: a communicator among engines[0..n) (rank r = engines[r]; ncclCommInitAll + peer
 * windows).  Engines that share a device are accepted too — several shards of one run on one GPU, exchanged through the
 * peer windows only (NCCL has no two ranks on one device): what the single-GPU tests use to exercise the exchange. */
int cpprob_sis_comm_init_local(cpprob_sis_engine * const * engines, int n_engines);

/* Single-process form of the same scheme: engines[r] (one per GPU, created by the caller with ONE seed) is rank r; the
 * calling thread queues every shard, the exchange and the merge, and waits once.  The local communicator is
 * made on first use (cpprob_sis_comm_init_local) and kept.  Estimators only (no trace emission).  Results are owned by
 * engines[0] and are bit-identical to a single-GPU run of the same seed. */
int cpprob_sis_run_multi(cpprob_sis_engine * const * engines, int n_engines, int model_id, const double * obs,
                         size_t n_obs, uint64_t n_particles, cpprob_sis_stats * out);

/* Writes <prefix>.ids and <prefix>.stats for the engine's last run (what infer_to_files does with EMIT_NONE). */
int cpprob_sis_write_summary(cpprob_sis_engine * e, const char * prefix, const cpprob_sis_stats * stats);

/* gathered: DEVICE pointer to [n_chunks_total][n_cols].  Returns 1 (not an error) if the weights
 * have to be re-based: call run_shard again with *m_ref_override = out->max_log_w. */
int cpprob_sis_merge(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                     const double * gathered, uint32_t n_chunks_total, int n_cols, double m_ref,
                     uint64_t n_particles_total, cpprob_sis_stats * out);

/* Same, on the raw output of an all-gather: `gathered` holds `world` segments of rows_per_rank rows, segment r
 * starting with rank r's rows (cpprob_sis_plan_rows gives how many; the rest of a segment is ignored).  The buffer
 * a shard returns always has room for one row more than it holds, so a rank may contribute
 * rows_per_rank = max over ranks of n_rows_local rows straight from cpprob_sis_partials.device_ptr.  The rows are
 * compacted on the device and merged exactly as cpprob_sis_merge does: same bits. */
int cpprob_sis_merge_padded(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * gathered,
                            int world, uint32_t rows_per_rank, int rows_per_chunk, int n_cols, double m_ref,
                            uint64_t n_particles_total, cpprob_sis_stats * out);

/* ---- replay ------------------------------------------------------------------------------------
 * Recomputes log_w for n recorded traces (host SoA rows as in cpprob_sis_block) by re-running the
 * model with every sample statement returning the recorded value.  Parity gate for
 * cpprob::observe / logpdf<> (cpprob.hpp:87-89): must match the reference's log-weights to 1e-12. */
int cpprob_sis_replay(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs,
                      const double * real_rows, const int32_t * int_rows, uint64_t stride, uint64_t n,
                      double * logw_out);

/* Estimators from externally supplied records (e.g. parsed posterior files): the device twin of
 * StatsPrinter::load_distr + operator<< (stats_printer.hpp:42-120). */
int cpprob_sis_reduce_records(cpprob_sis_engine * e, const double * real_rows, int n_real,
                              const int32_t * int_rows, int n_int, const double * log_w,
                              uint64_t stride, uint64_t n, cpprob_sis_stats * out);

/* ---- device distribution layer, exposed for testing and for callers that only need densities ----
 * kind: see CPPROB_SIS_DIST_*; params: up to 8 doubles (mean,sigma | a,b | min,max | p0..p7 |
 * mean | alpha(shape),beta(scale) | alpha,beta). */
enum {
    CPPROB_SIS_DIST_NORMAL = 0, CPPROB_SIS_DIST_UNIFORM_REAL = 1, CPPROB_SIS_DIST_UNIFORM_SMALLINT = 2,
    CPPROB_SIS_DIST_DISCRETE = 3, CPPROB_SIS_DIST_POISSON = 4, CPPROB_SIS_DIST_GAMMA = 5,
    CPPROB_SIS_DIST_BETA = 6
};
int cpprob_sis_logpdf(cpprob_sis_engine * e, int kind, const double * params, int n_params,
                      const double * x, uint64_t n, double * out);
int cpprob_sis_sample(cpprob_sis_engine * e, int kind, const double * params, int n_params,
                      uint64_t seed, uint64_t first_particle, uint64_t n, double * out);
/* raw Philox4x32-10 blocks: out[4*i..] = block(counter = ctr[4*i..], key = key[2*i..]) */
int cpprob_sis_philox(cpprob_sis_engine * e, const uint32_t * ctr, const uint32_t * key, uint64_t n, uint32_t * out);
/* fp64 elementary functions of include/cpprob/math/dmath.hpp: 0 log_unit 1 exp_weight 2 sin2pi 3 cos2pi (joint) 4 sqrt_pos 5 Box-Muller z0 from (u1,u2) packed as x[2i],x[2i+1] 6 log 7 cos_2pi 8 sin_2pi (single chain) 9 exp_weight_tab (table-assisted exp of the fused kernel; finite x, normal result) */
int cpprob_sis_dmath(cpprob_sis_engine * e, int fn, const double * x, uint64_t n, double * out);

/* ---- roofline denominators ----------------------------------------------------------------------*/
int cpprob_sis_measure_dfma_peak(cpprob_sis_engine * e, double * tflops, double * sm_clock_mhz_est);
int cpprob_sis_measure_store_peak(cpprob_sis_engine * e, double * gbytes_per_s);
/* issue-model probe: time of a fixed DFMA workload with `int_per_dfma` (0..3) ALU instructions interleaved per DFMA */
int cpprob_sis_probe_issue(cpprob_sis_engine * e, int int_per_dfma, double * ms_out);
/* latency probe: `chains` (1,2,4,8) independent DFMA chains per thread, blocks_per_sm CTAs of 256 threads per SM */
int cpprob_sis_probe_dfma_chains(cpprob_sis_engine * e, int chains, int blocks_per_sm, double * ms_out, double * dfma_per_thread);

#ifdef __cplusplus
}
#endif
#endif /* CPPROB_SIS_H */
