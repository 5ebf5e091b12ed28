// cpprob-b200: engine core behind the C ABI of include/cpprob_sis.h.
//
// Model-agnostic: all model code is reached through cpprob_sis_model_vtable launchers
// (model_vtable.cuh).  One engine drives one GPU with a compute stream and a copy stream; device
// and pinned buffers grow on demand and are kept between calls.
#include <algorithm>
#include <array>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>            // types only: the library is opened at run time (nccl_api below), never linked

#include "cpprob_sis.h"
#include "dist_kernels.cuh"
#include "model_vtable.cuh"
#include "posterior_text.hpp"
#include "reduce_kernels.cuh"
#include "text_kernels.cuh"

using namespace cpprob;
using namespace cpprob::engine;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string & msg)
{
    g_last_error = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        const cudaError_t cu_try_err = (expr);                                                         \
        if (cu_try_err != cudaSuccess) {                                                               \
            return fail(CPPROB_SIS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(cu_try_err)); \
        }                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// registry
// ------------------------------------------------------------------------------------------------
std::vector<const cpprob_sis_model_vtable *> & registry()
{
    static std::vector<const cpprob_sis_model_vtable *> r;
    return r;
}
std::mutex & registry_mutex()
{
    static std::mutex m;
    return m;
}

template<class T>
struct device_buffer {
    T * ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        const cudaError_t err = cudaMalloc(reinterpret_cast<void **>(&ptr), n * sizeof(T));
        if (err == cudaSuccess) cap = n;
        return err;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

template<class T>
struct pinned_buffer {
    T * ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        const cudaError_t err = cudaMallocHost(reinterpret_cast<void **>(&ptr), n * sizeof(T));
        if (err == cudaSuccess) cap = n;
        return err;
    }
    void release()
    {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

// ------------------------------------------------------------------------------------------------
// NCCL, opened at run time.  libcpprob_sis.so does not link against it: a single-GPU user needs no NCCL at all, and
// a process that already has one loaded (torch.distributed ships its own libnccl.so.2) must share that one.
// ------------------------------------------------------------------------------------------------
struct nccl_api {
    void * handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char * (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;         // set when loading failed
};

const nccl_api & nccl()
{
    static const nccl_api api = [] {
        nccl_api a;
        const char * names[] = {std::getenv("CPPROB_SIS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char * n : names) {
            if (!n || !*n) continue;
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) {
            a.why = std::string("NCCL could not be opened (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return a;
        }
        bool ok = true;
        auto sym = [&](const char * name) { void * p = dlsym(a.handle, name); ok = ok && p != nullptr; return p; };
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) {
            a.why = "the NCCL library that was found lacks an entry point this engine needs";
            a.handle = nullptr;
        }
        return a;
    }();
    return api;
}

#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        const ncclResult_t nccl_try_res = (expr);                                                              \
        if (nccl_try_res != ncclSuccess) {                                                                     \
            return fail(CPPROB_SIS_ENCCL, std::string(#expr) + ": " + nccl().GetErrorString(nccl_try_res));    \
        }                                                                                                      \
    } while (0)

}  // namespace

struct cpprob_sis_engine {
    int device = 0;
    int sm_count = 0;
    int blocks_per_sm = 0;
    uint64_t seed = 0;
    uint64_t max_batch = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_merge_begin = nullptr, ev_merge_end = nullptr;
    cudaEvent_t ev_computed[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_batch_begin[2] = {nullptr, nullptr};

    device_buffer<double> d_obs, d_pilot, d_partials, d_super, d_warp_partials, d_w[2], d_logw[2], d_real[2], d_gather;
    device_buffer<int> d_int[2];
    device_buffer<unsigned> d_counter;
    device_buffer<int_extra> d_int_extra;
    // device-side text stage: per-record lengths, CTA sums / offsets, the text itself, slot descriptors, ambiguity list
    device_buffer<unsigned> d_text_len[2];                        // [kind]
    device_buffer<unsigned long long> d_text_bsum[2], d_text_meta;   // meta: [kind]{total bytes, #ambiguous}
    device_buffer<char> d_text[2][2];                             // [double buffer][kind: 0 real, 1 int]
    device_buffer<text_slot> d_text_slots[2];
    device_buffer<text_flag> d_text_flags;                        // [double buffer][kind][kMaxTextFlags]
    pinned_buffer<char> h_text[2][2];
    pinned_buffer<unsigned long long> h_text_meta[2];
    pinned_buffer<text_flag> h_text_flags;
    cudaEvent_t ev_text[2] = {nullptr, nullptr}, ev_copy_begin[2] = {nullptr, nullptr};
    unsigned long long text_force_every = 0;                      // CPPROB_SIS_TEXT_FORCE_AMBIGUOUS (test hook)
    double text_kernel_ms = 0.0, text_copy_ms = 0.0, text_write_s = 0.0;   // stage times of the last emitting run
    uint64_t text_bytes = 0, text_fixups = 0;
    pinned_buffer<double> h_real[2], h_logw[2], h_merged, h_pilot, h_obs;
    pinned_buffer<int> h_int[2];

    // communicator of a multi-GPU run (cpprob_sis_comm_init / _init_local); rank 0 of 1 without one
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    // peer window of the communicator's ranks (reduce_kernels.cuh, k_push_rows): the exchange of an inference goes through
    // it when every rank could map every other rank's window; otherwise (or with CPPROB_SIS_EXCHANGE=nccl) ncclAllGather
    struct peer_window {
        unsigned char * local = nullptr;                   // [kPeerFlagBytes][2][buffer_bytes]
        unsigned char * peer[kMaxMergeRanks] = {};         // peer[comm_rank] == local
        bool opened[kMaxMergeRanks] = {};                  // mapped with cudaIpcOpenMemHandle (to be closed)
        size_t buffer_bytes = 0;
        unsigned long long epoch = 0;
        bool ready = false;
    } pw;

    // results kept alive for the caller
    const cpprob_sis_model_vtable * probe_vt = nullptr;           // what `structure` was probed for (probe_structure)
    uint64_t probe_seed = 0;
    std::vector<double> probe_obs;
    model_structure structure;
    std::vector<const char *> id_ptrs;
    std::vector<cpprob_sis_slot> slots;
    std::vector<double> real_mean, real_var, int_prob, sums;
    std::vector<long long> int_map;
    uint64_t launches = 0;
};

namespace {

int use_device(cpprob_sis_engine * e)
{
    CU_TRY(cudaSetDevice(e->device));
    return 0;
}

const cpprob_sis_model_vtable * model_of(int id)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    if (id < 0 || static_cast<size_t>(id) >= registry().size()) return nullptr;
    return registry()[static_cast<size_t>(id)];
}

int probe_structure(cpprob_sis_engine * e, const cpprob_sis_model_vtable * vt, const double * obs, size_t n_obs)
{
    if (vt->n_scalar_obs >= 0 && n_obs != static_cast<size_t>(vt->n_scalar_obs)) {
        return fail(CPPROB_SIS_EINVAL, std::string("model ") + vt->name + " takes " + std::to_string(vt->n_scalar_obs) +
                                           " observations, got " + std::to_string(n_obs));
    }
    if (n_obs == 0) {
        // cpprob.hpp:182 static_assert: the model has to receive the observed values
        return fail(CPPROB_SIS_EINVAL, "The function has to receive the observed values as parameters.");
    }
    // the structure is a function of (model, observations, seed): a repeated call with the same three — the usual case, an
    // inference is run again and again on one data set — keeps what the last probe found (hmm<1000>: a thousand host steps)
    if (e->probe_vt == vt && e->probe_seed == e->seed && e->probe_obs.size() == n_obs &&
        std::memcmp(e->probe_obs.data(), obs, n_obs * sizeof(double)) == 0) {
        return 0;
    }
    e->probe_vt = nullptr;
    e->structure = model_structure();
    vt->probe(obs, static_cast<int>(n_obs), e->seed, &e->structure);
    e->id_ptrs.clear();
    for (const auto & s : e->structure.ids) e->id_ptrs.push_back(s.c_str());
    e->slots.clear();
    for (const auto & s : e->structure.slots) {
        e->slots.push_back(cpprob_sis_slot{s.is_int ? 1 : 0, static_cast<int>(s.id), static_cast<int>(s.k), static_cast<int>(s.row), static_cast<int>(s.width)});
    }
    e->probe_vt = vt;
    e->probe_seed = e->seed;
    e->probe_obs.assign(obs, obs + n_obs);
    return 0;
}

// The observations go to the device through a pinned staging buffer (the caller's array is pageable: the runtime would
// stage it itself, more slowly).  Every call uploads them: they are the inference's input.
cudaError_t upload_obs(cpprob_sis_engine * e, const double * obs, size_t n_obs)
{
    if (cudaError_t err = e->h_obs.reserve(n_obs)) return err;
    if (cudaError_t err = cudaStreamSynchronize(e->compute)) return err;     // (idle between inferences: the last upload has been read)
    std::memcpy(e->h_obs.ptr, obs, n_obs * sizeof(double));
    return cudaMemcpyAsync(e->d_obs.ptr, e->h_obs.ptr, n_obs * sizeof(double), cudaMemcpyHostToDevice, e->compute);
}

struct shard_plan {
    uint32_t n_chunks_total = 0, chunk_first = 0, n_chunks_local = 0;
    uint64_t first_particle = 0, n_local = 0;
    // the rows a shard hands to the gather: partial sums of super-chunks = `super` consecutive chunks
    uint32_t super = 1, n_super_total = 0, super_first = 0, n_super_local = 0;
};

// Upper bound of the partial rows of a whole run (what is gathered and merged).  Test hook:
// CPPROB_SIS_MAX_PARTIAL_ROWS, read once per process.
uint32_t max_partial_rows()
{
    static const uint32_t v = [] {
        const char * s = std::getenv("CPPROB_SIS_MAX_PARTIAL_ROWS");
        const unsigned long x = s ? std::strtoul(s, nullptr, 10) : 4096ul;
        return static_cast<uint32_t>(std::max<unsigned long>(1ul, std::min<unsigned long>(x, 1ul << 20)));
    }();
    return v;
}

// The C chunks of a run are grouped into super-chunks of S = 2^k chunks, S the smallest that leaves at most
// max_partial_rows() of them: S depends on the run's size only, never on the GPU count.  A rank owns whole
// super-chunks and reduces each to ONE partial row (its chunk rows added in chunk order) before the gather, so the
// exchange is <= 4096 rows whatever the particle count, and the merged sums are still bit-identical for any world
// size.  Up to 4096 chunks (1.3e8 particles) S = 1 and a rank owns plain chunks.
shard_plan plan_shard(uint64_t n_total, int rank, int world)
{
    shard_plan p;
    p.n_chunks_total = static_cast<uint32_t>((n_total + kChunk - 1) / kChunk);
    while ((p.n_chunks_total + p.super - 1) / p.super > max_partial_rows()) p.super <<= 1;
    p.n_super_total = (p.n_chunks_total + p.super - 1) / p.super;
    const uint64_t s0 = static_cast<uint64_t>(p.n_super_total) * static_cast<uint64_t>(rank) / static_cast<uint64_t>(world);
    const uint64_t s1 = static_cast<uint64_t>(p.n_super_total) * static_cast<uint64_t>(rank + 1) / static_cast<uint64_t>(world);
    p.super_first = static_cast<uint32_t>(s0);
    p.n_super_local = static_cast<uint32_t>(s1 - s0);
    const uint64_t c0 = std::min<uint64_t>(p.n_chunks_total, s0 * p.super);
    const uint64_t c1 = std::min<uint64_t>(p.n_chunks_total, s1 * p.super);
    p.chunk_first = static_cast<uint32_t>(c0);
    p.n_chunks_local = static_cast<uint32_t>(c1 - c0);
    p.first_particle = c0 * kChunk;
    const uint64_t end = std::min<uint64_t>(n_total, c1 * kChunk);
    p.n_local = end > p.first_particle ? end - p.first_particle : 0;
    return p;
}

struct hist_window {
    long long lo = 0;
    int bins = 0;
};

cudaError_t launch_rows_moments(cudaStream_t s, unsigned n_sub_chunks, const double * rows, const double * w, unsigned long long stride,
                                unsigned long long n, int n_real, double * partials, int n_cols)
{
    // 70 KB of staging areas per CTA: beyond the default limit (the attribute is per device, and cheap to set)
    if (cudaError_t err = cudaFuncSetAttribute(k_rows_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kRowsMomentsSmem))) return err;
    const unsigned grid = n_sub_chunks * static_cast<unsigned>((n_real + 31) / 32);
    k_rows_moments<<<grid, kBlock, kRowsMomentsSmem, s>>>(rows, w, stride, n, n_real, partials, n_cols);
    return cudaGetLastError();
}

// all histogram passes for one batch (kRowsHistBins bins per launch); counts the launches in *launches
cudaError_t launch_hist_all(cudaStream_t s, unsigned n_chunks, int n_int, const int * rows, const double * w,
                            unsigned long long stride, unsigned long long n, hist_window hw, int col0,
                            double * partials, int n_cols, uint64_t * launches)
{
    const unsigned grid = n_chunks * static_cast<unsigned>((n_int + 31) / 32);        // one CTA per (sub-chunk, group of 32 rows)
    for (int off = 0; off < hw.bins; off += kRowsHistBins) {
        const int here = std::min(kRowsHistBins, hw.bins - off);
        k_rows_hist<<<grid, kBlock, 0, s>>>(rows, w, stride, n, n_int, hw.lo + off, off, here, hw.bins, col0, partials, n_cols);
        if (cudaError_t err = cudaGetLastError()) return err;
        ++*launches;
    }
    return cudaSuccess;
}

// Runs the pilot and folds its per-tile maxima on the host.  pilot[0] = m_ref (finite; 0 when every pilot
// weight is -inf / nan), pilot[1] = min int, pilot[2] = max int (min > max when there are no int predicts).
// Leaves m_ref in e->d_pilot[0] for the kernels.
int run_pilot(cpprob_sis_engine * e, const cpprob_sis_model_vtable * vt, const philox_keys & keys, size_t n_obs, uint64_t n_total,
              const double * m_ref_override, double pilot[3])
{
    const int n_pilot = static_cast<int>(std::min<uint64_t>(n_total, kPilot));
    const int tiles = (n_pilot + 511) / 512;
    double raw[3 * kPilotTiles];
    CU_TRY(e->d_pilot.reserve(3 * kPilotTiles));
    CU_TRY(vt->launch_pilot(e->compute, &keys, e->d_obs.ptr, static_cast<int>(n_obs), n_pilot, e->d_pilot.ptr, nullptr));
    CU_TRY(cudaMemcpyAsync(raw, e->d_pilot.ptr, 3 * tiles * sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    double mx = -std::numeric_limits<double>::infinity(), nmin = mx, imax = mx;
    for (int t = 0; t < tiles; ++t) {
        mx = std::fmax(mx, raw[3 * t]);
        nmin = std::fmax(nmin, raw[3 * t + 1]);
        imax = std::fmax(imax, raw[3 * t + 2]);
    }
    pilot[0] = (mx > -1.0e300 && mx < 1.0e300) ? mx : 0.0;
    pilot[1] = -nmin;
    pilot[2] = imax;
    const double m_ref = m_ref_override ? *m_ref_override : pilot[0];
    CU_TRY(cudaMemcpyAsync(e->d_pilot.ptr, &m_ref, sizeof(double), cudaMemcpyHostToDevice, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));   // m_ref lives on this stack frame
    pilot[0] = m_ref;
    return 0;
}

constexpr unsigned kMaxTextFlags = 4096;   // ambiguous records per batch and kind the host can re-format

struct shard_options {
    int emit = CPPROB_SIS_EMIT_NONE;
    int force_rows = 0;
    cpprob_sis_block_fn on_block = nullptr;
    void * user = nullptr;
    cpprob::text::posterior_writer * text_writer = nullptr;   // EMIT_ALL with the records formatted on the GPU
    bool no_wait = false;    // fused path only: return right after the launches; merge_impl(..., pending) collects m_ref and the time
    bool count_text_only = false;   // with text_writer: only measure the text of this shard (shard_result::text_total), write nothing
    bool push_peers = false;        // multi-GPU: hand the rows on through the ranks' peer windows (if there are any and the rows fit)
};

constexpr int kMinStagedWarps = 4;          // fewer resident warps than this: the row path is the better choice

// test / A-B hook: CPPROB_SIS_STAGED=0 sends estimator-only runs of long traces through the row path, as before
bool staged_enabled()
{
    static const bool on = [] {
        const char * s = std::getenv("CPPROB_SIS_STAGED");
        return !(s && std::strcmp(s, "0") == 0);
    }();
    return on;
}

struct shard_result {
    int path = CPPROB_SIS_PATH_ROWS;
    bool pushed = false;                                           // the rows are in every rank's peer window (epoch / offset below)
    unsigned long long push_epoch = 0, push_offset = 0;
    unsigned long long text_total[2] = {0, 0};                     // count_text_only: bytes of this shard's .real / .int text
    shard_plan plan;
    uint32_t rows_per_chunk = 1;       // partial rows per kChunk particles: 1 (fused) or kChunk / kSubChunk (row path)
    uint32_t n_rows_local = 0, n_rows_total = 0, row_first = 0;   // the rows handed on (super-chunk rows, see plan_shard)
    const double * rows = nullptr;                                // [n_rows_local][n_cols] on the device
    bool waiting = false;                                         // launched with no_wait: m_ref / device_ms not collected yet
    int n_cols = 0;
    hist_window hw;
    double m_ref = 0.0;
    double device_ms = 0.0;
    uint64_t launches = 0;
};

// the rows every rank owns (host arithmetic, identical on every rank) as the layout of the gathered buffer
int make_gather_layout(uint64_t n_total, int world, int rows_per_chunk, gather_layout * lay)
{
    if (world > kMaxMergeRanks) return fail(CPPROB_SIS_EINVAL, "more than 64 ranks");
    uint32_t seen = 0, most = 0, total = 0;
    for (int r = 0; r < world; ++r) {
        uint32_t first = 0, n_local = 0;
        if (int rc = cpprob_sis_plan_rows(n_total, r, world, rows_per_chunk, &first, &n_local, &total)) return rc;
        if (first != seen) return fail(CPPROB_SIS_EINVAL, "inconsistent row plan");
        lay->first[r] = first;
        seen += n_local;
        most = std::max(most, n_local);
    }
    lay->first[world] = seen;
    lay->world = static_cast<unsigned>(world);
    lay->rows_per_rank = std::max<uint32_t>(most, 1u);
    return 0;
}

// The particle pass of one rank.  Leaves [n_chunks_local][n_cols] partial sums in e->d_partials.
int run_shard_impl(cpprob_sis_engine * e, const cpprob_sis_model_vtable * vt, const double * obs, size_t n_obs,
                   uint64_t n_total, int rank, int world, const double * m_ref_override, const hist_window * hw_override,
                   const shard_options & opt, shard_result * res)
{
    if (world <= 0 || rank < 0 || rank >= world) return fail(CPPROB_SIS_EINVAL, "bad rank / world");
    if (n_total == 0) return fail(CPPROB_SIS_EINVAL, "n_particles must be positive");
    if (n_total > (static_cast<uint64_t>(1) << 46)) return fail(CPPROB_SIS_EINVAL, "n_particles too large");
    if (int rc = use_device(e)) return rc;
    if (int rc = probe_structure(e, vt, obs, n_obs)) return rc;
    const int n_real = static_cast<int>(e->structure.n_real);
    const int n_int = static_cast<int>(e->structure.n_int);
    const shard_plan plan = plan_shard(n_total, rank, world);
    if (plan.n_chunks_local > (1u << 27)) {      // the fused kernel counts (chunk, part <= 4, warp slot) units in 32 bits
        return fail(CPPROB_SIS_EINVAL, "more than 2^42 (4.4e12) particles per GPU in one call");
    }
    res->plan = plan;
    res->launches = 0;
    res->device_ms = 0.0;

    CU_TRY(e->d_obs.reserve(n_obs));
    CU_TRY(e->d_pilot.reserve(3 * kPilotTiles));
    if (e->d_counter.cap < 2) {                  // [unit counter of the particle kernel][CTA-done counter of the pilot], zero between launches
        CU_TRY(e->d_counter.reserve(2));
        CU_TRY(cudaMemsetAsync(e->d_counter.ptr, 0, 2 * sizeof(unsigned), e->compute));
    }
    CU_TRY(upload_obs(e, obs, n_obs));

    // pilot: m_ref and the int window, identical on every rank.  The device-timed region of a run
    // starts here: it covers the pilot, the particle kernel(s) and the row reductions.
    const philox_keys keys(e->seed);
    CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
    double pilot[3] = {0, 0, 0};
    // A model without int predicts needs nothing from the pilot on the host before its kernels start: the maxima are
    // folded on the device and m_ref is read back with the results (one host round trip less per inference).
    // Paths: estimator-only runs keep the trace off HBM — in registers when the model has at most
    // kMaxFusedReal real predicts and no int predicts (k_sis_fused), else in per-warp shared-memory staging areas
    // (k_sis_staged) as long as those fit; emitting runs, forced runs and very long traces write SoA rows (k_sis_rows).
    const bool est_only = !opt.force_rows && opt.emit == CPPROB_SIS_EMIT_NONE;
    const bool fused = est_only && n_int == 0 && n_real <= kMaxFusedReal;
    int staged_warps = 0;
    // a model that declares the range of its int predicts ([0, int_states)) needs no pilot for the histogram window
    const int declared = (n_int > 0 && !hw_override) ? vt->int_states : 0;
    if (est_only && !fused && staged_enabled()) {
        if (n_int == 0) staged_warps = vt->staged_warps(static_cast<int>(n_obs), n_real, 0, 0);
        else if (declared > 0) staged_warps = vt->staged_warps(static_cast<int>(n_obs), n_real, n_int, declared);
    }
    const bool fused_early = fused || staged_warps >= kMinStagedWarps;
    if (fused_early) {
        const int n_pilot = static_cast<int>(std::min<uint64_t>(n_total, kPilot));
        // Nothing but kernels between here and the end of the particle pass: the pilot's last CTA folds the tiles into m_ref
        // and zeroes the unit counter, and m_ref is read back with the results (a second launch, a copy and a memset in the
        // middle of the chain cost several microseconds each, which is what a short run — one GPU's share of a
        // strong-scaling run — is made of).
        CU_TRY(e->h_pilot.reserve(1));
        const pilot_finalize fin = {e->d_counter.ptr, m_ref_override ? 1 : 0, m_ref_override ? *m_ref_override : 0.0};
        CU_TRY(vt->launch_pilot(e->compute, &keys, e->d_obs.ptr, static_cast<int>(n_obs), n_pilot, e->d_pilot.ptr, &fin));
        res->launches += 1;
    } else {
        if (int rc = run_pilot(e, vt, keys, n_obs, n_total, m_ref_override, pilot)) return rc;
        ++res->launches;
    }
    const double m_ref = pilot[0];             // fused_early: filled in after the particle kernel's sync
    hist_window hw;
    if (n_int > 0) {
        if (hw_override) {
            hw = *hw_override;
        } else if (declared > 0) {                 // the model says where its int predicts lie: same window on every path
            hw.lo = 0;
            hw.bins = declared;
        } else if (pilot[1] <= pilot[2]) {
            hw.lo = static_cast<long long>(pilot[1]);
            hw.bins = static_cast<int>(std::min<double>(pilot[2] - pilot[1] + 1.0, 1.0e9));
        } else {
            hw.lo = 0;
            hw.bins = 1;
        }
        if (hw.bins > 4096) return fail(CPPROB_SIS_ERANGE, "int predicts span more than 4096 values");
    }
    res->hw = hw;
    res->m_ref = m_ref;
    const int n_cols = kBaseCols + 2 * n_real + n_int * hw.bins;
    res->n_cols = n_cols;
    // (a model compiled for four states per staged byte only fits the window it declared)
    const bool packed_model = vt->int_states >= 1 && vt->int_states <= 4;
    if (est_only && !fused && n_int > 0 && staged_enabled() && !fused_early && (!packed_model || (hw.lo == 0 && hw.bins == vt->int_states))) {
        staged_warps = vt->staged_warps(static_cast<int>(n_obs), n_real, n_int, hw.bins);
    }
    const bool staged = !fused && staged_warps >= kMinStagedWarps;
    res->path = fused ? CPPROB_SIS_PATH_FUSED : (staged ? CPPROB_SIS_PATH_STAGED : CPPROB_SIS_PATH_ROWS);
    // partial-sum rows: one per chunk on the fused path, one per sub-chunk on the row path (same on every rank)
    const unsigned row_particles = fused ? kChunk : kSubChunk;
    res->rows_per_chunk = kChunk / row_particles;
    const uint32_t kernel_rows = static_cast<uint32_t>((plan.n_local + row_particles - 1) / row_particles);   // rows the kernels write
    const uint32_t rows_per_super = plan.super * res->rows_per_chunk;
    res->n_rows_total = plan.n_super_total;
    res->row_first = plan.super_first;
    res->n_rows_local = plan.n_super_local;
    if (rows_per_super == 1) {                 // small run on the fused path: the kernel rows are the rows handed on
        res->n_rows_total = plan.n_chunks_total;
        res->row_first = plan.chunk_first;
        res->n_rows_local = kernel_rows;
    }
    res->rows = nullptr;
    // Multi-GPU over peer memory: the kernel that produces the rows this rank hands on also stores them into every rank's
    // gather buffer and publishes the inference's epoch there (reduce_kernels.cuh).  Whether that happens is a function of
    // the communicator and the run's shape only, so every rank decides alike.
    peer_push pp;
    std::memset(&pp, 0, sizeof pp);
    res->pushed = false;
    if (opt.push_peers && world > 1 && e->pw.ready && world == e->comm_world && rank == e->comm_rank) {
        gather_layout lay;
        if (int rc = make_gather_layout(n_total, world, static_cast<int>(res->rows_per_chunk), &lay)) return rc;
        const size_t count = static_cast<size_t>(lay.rows_per_rank) * n_cols;
        if (count * static_cast<size_t>(world) * sizeof(double) <= e->pw.buffer_bytes) {
            pp.epoch = ++e->pw.epoch;
            pp.buffer_offset_bytes = kPeerFlagBytes + (pp.epoch & 1ull) * e->pw.buffer_bytes;
            pp.segment_doubles = count;
            for (int p = 0; p < world; ++p) pp.t.window[p] = e->pw.peer[p];
            pp.t.world = static_cast<unsigned>(world);
            pp.t.rank = static_cast<unsigned>(rank);
            res->pushed = true;
            res->push_epoch = pp.epoch;
            res->push_offset = pp.buffer_offset_bytes;
        }
    }
    // m_ref to the host, queued behind the kernels (a caller that goes on to the merge gets it from there instead)
    auto fetch_m_ref = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(e->h_pilot.ptr, e->d_pilot.ptr, sizeof(double), cudaMemcpyDeviceToHost, e->compute));
        return 0;
    };
    if (plan.n_chunks_local == 0) {
        if (res->pushed) {                     // a rank without particles still publishes its (empty) segment
            k_push_rows<<<1, kBlock, 0, e->compute>>>(nullptr, 0ull, pp);
            CU_TRY(cudaGetLastError());
            ++res->launches;
        }
        if (fused_early) {                     // and reports the run's m_ref
            if (int rc = fetch_m_ref()) return rc;
            CU_TRY(cudaStreamSynchronize(e->compute));
            res->m_ref = e->h_pilot.ptr[0];
        }
        return 0;
    }
    CU_TRY(e->d_partials.reserve(static_cast<size_t>(kernel_rows + 1) * n_cols));   // + 1: see cpprob_sis_merge_padded
    res->rows = e->d_partials.ptr;
    // kernel rows -> super-chunk rows (in row order), on the compute stream; then, with peers, into their windows
    auto fold_rows = [&]() -> int {
        if (rows_per_super > 1) {
            CU_TRY(e->d_super.reserve(static_cast<size_t>(plan.n_super_local + 1) * n_cols));
            const unsigned long long n_out = static_cast<unsigned long long>(plan.n_super_local) * n_cols;
            k_fold_rows<<<static_cast<unsigned>((n_out + kBlock - 1) / kBlock), kBlock, 0, e->compute>>>(
                e->d_partials.ptr, kernel_rows, rows_per_super, n_cols, kMaxColsMask, e->d_super.ptr, plan.n_super_local);
            CU_TRY(cudaGetLastError());
            ++res->launches;
            res->rows = e->d_super.ptr;
        }
        if (res->pushed) {
            const unsigned long long n_doubles = static_cast<unsigned long long>(res->n_rows_local) * n_cols;
            const unsigned grid = static_cast<unsigned>(std::min<unsigned long long>(std::max<unsigned long long>((n_doubles + kBlock - 1) / kBlock, 1ull), 64ull));
            k_push_rows<<<grid, kBlock, 0, e->compute>>>(res->rows, n_doubles, pp);
            CU_TRY(cudaGetLastError());
            ++res->launches;
        }
        return 0;
    };
    // the particle kernels' unit rows (`per_unit` per kernel row, nv columns each) -> the rows handed on (-> the peers), one launch
    auto fold_units = [&](int per_unit, int nv) -> int {
        double * out = e->d_partials.ptr;
        unsigned n_out_rows = kernel_rows;
        if (rows_per_super > 1) {
            CU_TRY(e->d_super.reserve(static_cast<size_t>(plan.n_super_local + 1) * n_cols));
            out = e->d_super.ptr;
            n_out_rows = plan.n_super_local;
        }
        const unsigned long long n_out = static_cast<unsigned long long>(n_out_rows) * n_cols;
        unsigned group = 1;                          // lanes per output element: the power of two >= rows_per_super, at most a warp
        while (group < rows_per_super && group < 32u) group <<= 1;
        const unsigned long long threads = n_out * group;
        k_fold_units<<<static_cast<unsigned>((threads + kBlock - 1) / kBlock), kBlock, 0, e->compute>>>(
            e->d_warp_partials.ptr, kernel_rows, per_unit, nv, rows_per_super, n_cols, kMaxColsMask, out, n_out_rows, group, pp);
        CU_TRY(cudaGetLastError());
        ++res->launches;
        res->rows = out;
        return 0;
    };

    run_args a;
    std::memset(&a, 0, sizeof a);
    a.keys = keys;
    a.n_obs = static_cast<int>(n_obs);
    a.obs = e->d_obs.ptr;
    a.m_ref = e->d_pilot.ptr;
    a.chunk_counter = e->d_counter.ptr;
    a.n_cols = n_cols;
    a.hist_lo = hw.lo;
    a.hist_bins = hw.bins;

    if (staged) {
        // one CTA per SM, as many warps as the staging areas allow; any warp takes any (sub-chunk, slot) unit
        const uint64_t units = static_cast<uint64_t>(kernel_rows) * kSlotsPerChunk;
        const int grid = static_cast<int>(std::min<uint64_t>((units + staged_warps - 1) / staged_warps, static_cast<uint64_t>(e->sm_count)));
        a.first_particle = plan.first_particle;
        a.n_particles = plan.n_local;
        a.n_chunks = kernel_rows;
        a.partials = e->d_partials.ptr;
        a.n_real = n_real;
        a.n_int = n_int;
        CU_TRY(e->d_warp_partials.reserve(static_cast<size_t>(units) * n_cols));
        a.warp_partials = e->d_warp_partials.ptr;
        if (!fused_early) CU_TRY(cudaMemsetAsync(e->d_counter.ptr, 0, sizeof(unsigned), e->compute));   // else: zeroed by the pilot's last CTA
        CU_TRY(vt->launch_staged(e->compute, grid, staged_warps, &a));
        ++res->launches;
        if (int rc = fold_units(kSlotsPerChunk, n_cols)) return rc;
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        if (opt.no_wait && fused_early) {          // m_ref still on the device: collected after the merge
            res->waiting = true;
            return 0;
        }
        if (fused_early) { if (int rc = fetch_m_ref()) return rc; }
        CU_TRY(cudaStreamSynchronize(e->compute));
        if (fused_early) res->m_ref = e->h_pilot.ptr[0];
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        res->device_ms = ms;
        return 0;
    }

    if (fused) {
        const int nr = n_real <= 1 ? 1 : (n_real == 2 ? 2 : 4);
        int occ = vt->occupancy(nr == 1 ? 0 : (nr == 2 ? 1 : 2), static_cast<int>(n_obs));
        if (occ <= 0) occ = 1;
        if (e->blocks_per_sm > 0) occ = std::min(occ, e->blocks_per_sm);
        // warp-autonomous kernel: enough CTAs to give every (chunk, warp slot) unit a warp, at most the resident set
        const uint64_t ctas_needed = (static_cast<uint64_t>(plan.n_chunks_local) * kFusedRowsPerChunk * 32 + fused_block(nr) - 1) / fused_block(nr);
        const int grid = static_cast<int>(std::min<uint64_t>(ctas_needed, static_cast<uint64_t>(e->sm_count) * occ));
        a.first_particle = plan.first_particle;
        a.n_particles = plan.n_local;
        a.n_chunks = plan.n_chunks_local;
        a.partials = e->d_partials.ptr;
        const int nv = kBaseCols + 2 * nr;
        CU_TRY(e->d_warp_partials.reserve(static_cast<size_t>(plan.n_chunks_local) * kFusedRowsPerChunk * nv));
        a.warp_partials = e->d_warp_partials.ptr;
        CU_TRY(vt->launch_fused(e->compute, grid, nr, &a));      // (unit counter: zeroed by the pilot's last CTA)
        ++res->launches;
        if (int rc = fold_units(kFusedRowsPerChunk, nv)) return rc;
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        if (opt.no_wait) {
            res->waiting = true;
            return 0;
        }
        if (int rc = fetch_m_ref()) return rc;
        CU_TRY(cudaStreamSynchronize(e->compute));
        res->m_ref = e->h_pilot.ptr[0];
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        res->device_ms = ms;
        return 0;
    }

    // ---- row path: batches of whole chunks ------------------------------------------------------
    const bool emit = opt.emit == CPPROB_SIS_EMIT_ALL;
    const uint64_t bytes_per_particle = 8ull * n_real + 4ull * n_int + 16ull;
    // EMIT_ALL with a text writer: the records are formatted on the GPU (text_kernels.cuh) and only text crosses
    // PCIe; a line is ~3.3x its binary record, so the batches are smaller
    const bool text_mode = emit && opt.text_writer != nullptr;
    const uint64_t budget = text_mode ? (128ull << 20) : (emit ? (512ull << 20) : (8192ull << 20));
    uint64_t cap = e->max_batch ? e->max_batch : budget / bytes_per_particle;
    cap = std::max<uint64_t>(kChunk, cap / kChunk * kChunk);
    cap = std::min<uint64_t>(cap, static_cast<uint64_t>(plan.n_chunks_local) * kChunk);
    const int n_buf = emit ? 2 : 1;
    for (int b = 0; b < n_buf; ++b) {
        CU_TRY(e->d_real[b].reserve(std::max<size_t>(1, static_cast<size_t>(n_real) * cap)));
        CU_TRY(e->d_int[b].reserve(std::max<size_t>(1, static_cast<size_t>(n_int) * cap)));
        CU_TRY(e->d_logw[b].reserve(cap));
        CU_TRY(e->d_w[b].reserve(cap));
        if (emit && !text_mode) {
            CU_TRY(e->h_real[b].reserve(std::max<size_t>(1, static_cast<size_t>(n_real) * cap)));
            CU_TRY(e->h_int[b].reserve(std::max<size_t>(1, static_cast<size_t>(n_int) * cap)));
            CU_TRY(e->h_logw[b].reserve(cap));
        }
    }
    if (n_int > 0) CU_TRY(e->d_int_extra.reserve(static_cast<size_t>(cap / kSubChunk)));
    int occ = vt->occupancy(3, static_cast<int>(n_obs));
    if (occ <= 0) occ = 1;
    if (e->blocks_per_sm > 0) occ = std::min(occ, e->blocks_per_sm);

    // text stage set-up: slot descriptors and per-record length / CTA offset arrays on the device
    int n_text_slots[2] = {0, 0};
    if (text_mode) {
        e->text_kernel_ms = e->text_copy_ms = e->text_write_s = 0.0;
        e->text_bytes = e->text_fixups = 0;
        for (int kind = 0; kind < 2; ++kind) {
            const std::vector<cpprob_sis_slot> & sl = opt.text_writer->slots(kind == 1);
            if (sl.empty()) continue;
            std::vector<text_slot> h;
            for (const auto & s : sl) h.push_back(text_slot{s.id, s.row, s.width});
            n_text_slots[kind] = static_cast<int>(h.size());
            CU_TRY(e->d_text_slots[kind].reserve(h.size()));
            CU_TRY(cudaMemcpy(e->d_text_slots[kind].ptr, h.data(), h.size() * sizeof(text_slot), cudaMemcpyHostToDevice));
            CU_TRY(e->d_text_len[kind].reserve(cap));
            CU_TRY(e->d_text_bsum[kind].reserve((cap + kTextBlock - 1) / kTextBlock));
        }
        CU_TRY(e->d_text_meta.reserve(4));
        CU_TRY(e->d_text_flags.reserve(4 * kMaxTextFlags));       // [double buffer][kind]: batch b+1 must not overwrite the list batch b is still copying
        CU_TRY(e->h_text_flags.reserve(4 * kMaxTextFlags));
        for (int b = 0; b < 2; ++b) CU_TRY(e->h_text_meta[b].reserve(4));
    }

    const uint64_t n_batches = (plan.n_local + cap - 1) / cap;
    struct pending_block { bool valid = false; uint64_t first = 0, n = 0; unsigned long long text_bytes[2] = {0, 0}, text_flags[2] = {0, 0}; };
    pending_block pending[2];
    // a record the GPU formatter could not decide (text_format.cuh): its line is re-made on the host from the
    // values still in the device rows and patched into the pinned text
    auto fix_up = [&](int buf, int kind, const text_flag & fl) -> int {
        std::vector<double> rv(std::max(1, n_real));
        std::vector<int> iv(std::max(1, n_int));
        double lw = 0.0;
        if (n_real > 0) CU_TRY(cudaMemcpy2D(rv.data(), sizeof(double), e->d_real[buf].ptr + fl.record, cap * sizeof(double), sizeof(double), n_real, cudaMemcpyDeviceToHost));
        if (n_int > 0) CU_TRY(cudaMemcpy2D(iv.data(), sizeof(int), e->d_int[buf].ptr + fl.record, cap * sizeof(int), sizeof(int), n_int, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&lw, e->d_logw[buf].ptr + fl.record, sizeof(double), cudaMemcpyDeviceToHost));
        cpprob_sis_block one;
        one.first_particle = pending[buf].first + fl.record;
        one.n = 1;
        one.stride = 1;
        one.n_real = n_real;
        one.n_int = n_int;
        one.real_rows = rv.data();
        one.int_rows = iv.data();
        one.log_w = &lw;
        const std::string line = opt.text_writer->format_record(one, kind == 1, 0);
        if (line.size() != fl.length || fl.offset + fl.length > pending[buf].text_bytes[kind]) {
            return fail(CPPROB_SIS_EIO, "device text stage: host re-formatting of an ambiguous record changed its length");
        }
        std::memcpy(e->h_text[buf][kind].ptr + fl.offset, line.data(), line.size());
        ++e->text_fixups;
        return 0;
    };
    auto deliver = [&](int buf) -> int {
        if (!pending[buf].valid) return 0;
        CU_TRY(cudaEventSynchronize(e->ev_copied[buf]));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_batch_begin[buf], e->ev_computed[buf]));
        res->device_ms += ms;
        pending[buf].valid = false;
        if (text_mode) {
            CU_TRY(cudaEventElapsedTime(&ms, e->ev_copy_begin[buf], e->ev_copied[buf]));
            e->text_copy_ms += ms;
            const auto t0 = std::chrono::steady_clock::now();
            for (int kind = 0; kind < 2; ++kind) {
                if (!n_text_slots[kind]) continue;
                for (unsigned long long f = 0; f < pending[buf].text_flags[kind]; ++f) {
                    if (int rc = fix_up(buf, kind, e->h_text_flags.ptr[(2 * buf + kind) * kMaxTextFlags + f])) return rc;
                }
                if (!opt.text_writer->write_text(kind == 1, e->h_text[buf][kind].ptr, pending[buf].text_bytes[kind])) {
                    return fail(CPPROB_SIS_EIO, std::string("cannot write the posterior file: ") + std::strerror(errno));
                }
                e->text_bytes += pending[buf].text_bytes[kind];
            }
            e->text_write_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            return 0;
        }
        if (opt.on_block) {
            cpprob_sis_block blk;
            blk.first_particle = pending[buf].first;
            blk.n = pending[buf].n;
            blk.stride = cap;
            blk.n_real = n_real;
            blk.n_int = n_int;
            blk.real_rows = e->h_real[buf].ptr;
            blk.int_rows = e->h_int[buf].ptr;
            blk.log_w = e->h_logw[buf].ptr;
            if (opt.on_block(opt.user, &blk) != 0) return fail(CPPROB_SIS_EIO, "trace block consumer failed");
        }
        return 0;
    };

    for (uint64_t b = 0; b < n_batches; ++b) {
        const int buf = emit ? static_cast<int>(b & 1) : 0;
        const uint64_t off = b * cap;
        const uint64_t n_here = std::min<uint64_t>(cap, plan.n_local - off);
        const unsigned subs_here = static_cast<unsigned>((n_here + kSubChunk - 1) / kSubChunk);
        if (emit) {
            // the previous user of this buffer pair (batch b-2) must have been handed to the consumer
            if (int rc = deliver(buf)) return rc;
            // ... and its device->host copy must be done before the rows are overwritten
            if (b >= 2) CU_TRY(cudaStreamWaitEvent(e->compute, e->ev_copied[buf], 0));
            CU_TRY(cudaEventRecord(e->ev_batch_begin[buf], e->compute));
        }
        a.first_particle = plan.first_particle + off;
        a.n_particles = n_here;
        a.n_chunks = subs_here;
        a.chunk = kSubChunk;
        a.partials = e->d_partials.ptr + (off / kSubChunk) * n_cols;
        a.real_rows = e->d_real[buf].ptr;
        a.int_rows = e->d_int[buf].ptr;
        a.logw = e->d_logw[buf].ptr;
        a.w = e->d_w[buf].ptr;
        a.row_stride = cap;
        a.int_extras = n_int > 0 ? e->d_int_extra.ptr : nullptr;
        if (n_int > 0) {
            k_init_int_extra<<<(subs_here + kBlock - 1) / kBlock, kBlock, 0, e->compute>>>(e->d_int_extra.ptr, subs_here);
            CU_TRY(cudaGetLastError());
            ++res->launches;
        }
        const uint64_t n_tiles = (n_here + 2 * kPairStride - 1) / (2 * kPairStride);
        const int grid = static_cast<int>(std::min<uint64_t>(n_tiles, static_cast<uint64_t>(e->sm_count) * occ));
        CU_TRY(vt->launch_rows(e->compute, grid, &a));
        ++res->launches;
        k_row_base<<<subs_here, kBlock, 0, e->compute>>>(a.logw, n_here, kSubChunk, e->d_pilot.ptr, a.w, a.int_extras, hw.lo, hw.bins, a.partials, n_cols);
        CU_TRY(cudaGetLastError());
        ++res->launches;
        if (n_real > 0 && !opt.count_text_only) {
            CU_TRY(launch_rows_moments(e->compute, subs_here, a.real_rows, a.w, cap, n_here, n_real, a.partials, n_cols));
            ++res->launches;
        }
        if (n_int > 0 && !opt.count_text_only) {
            CU_TRY(launch_hist_all(e->compute, subs_here, n_int, a.int_rows, a.w, cap, n_here, hw, kBaseCols + 2 * n_real,
                                   a.partials, n_cols, &res->launches));
        }
        if (text_mode) {
            CU_TRY(cudaEventRecord(e->ev_computed[buf], e->compute));
            CU_TRY(cudaMemsetAsync(e->d_text_meta.ptr, 0, 4 * sizeof(unsigned long long), e->compute));
            const unsigned text_blocks = static_cast<unsigned>((n_here + kTextBlock - 1) / kTextBlock);
            text_args ta[2];
            // pass 1: every record's length, the CTA offsets and the batch's text size, per kind
            for (int kind = 0; kind < 2; ++kind) {
                if (!n_text_slots[kind]) continue;
                ta[kind].slots = e->d_text_slots[kind].ptr;
                ta[kind].n_slots = n_text_slots[kind];
                ta[kind].is_int = kind;
                ta[kind].real_rows = a.real_rows;
                ta[kind].int_rows = a.int_rows;
                ta[kind].logw = a.logw;
                ta[kind].stride = cap;
                ta[kind].n = n_here;
                ta[kind].first_particle = plan.first_particle + off;
                ta[kind].force_every = e->text_force_every;
                k_text_lengths<<<text_blocks, kTextBlock, 0, e->compute>>>(ta[kind], e->d_text_len[kind].ptr, e->d_text_bsum[kind].ptr);
                CU_TRY(cudaGetLastError());
                k_text_scan<<<1, 1024, 0, e->compute>>>(e->d_text_bsum[kind].ptr, text_blocks, e->d_text_meta.ptr + 2 * kind);
                CU_TRY(cudaGetLastError());
                res->launches += 2;
            }
            CU_TRY(cudaMemcpyAsync(e->h_text_meta[buf].ptr, e->d_text_meta.ptr, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->compute));
            CU_TRY(cudaStreamSynchronize(e->compute));
            if (opt.count_text_only) {           // multi-GPU emission, first pass: the size of this rank's text is all that is wanted
                for (int kind = 0; kind < 2; ++kind) if (n_text_slots[kind]) res->text_total[kind] += e->h_text_meta[buf].ptr[2 * kind];
                continue;
            }
            // the text buffers are sized from the measured totals (grow-only, 1/16 head room), not from a worst case:
            // an hmm<1000> line is 6 KB, its worst case 28 KB
            for (int kind = 0; kind < 2; ++kind) {
                if (!n_text_slots[kind]) continue;
                const size_t total = e->h_text_meta[buf].ptr[2 * kind];
                pending[buf].text_bytes[kind] = total;
                if (total > e->d_text[buf][kind].cap) CU_TRY(e->d_text[buf][kind].reserve(total + total / 16 + 4096));
                if (total > e->h_text[buf][kind].cap) CU_TRY(e->h_text[buf][kind].reserve(total + total / 16 + 4096));
            }
            // pass 2: every record formatted again at its final offset
            for (int kind = 0; kind < 2; ++kind) {
                if (!n_text_slots[kind]) continue;
                k_text_write<<<text_blocks, kTextBlock, 0, e->compute>>>(ta[kind], e->d_text_len[kind].ptr, e->d_text_bsum[kind].ptr,
                                                                         e->d_text[buf][kind].ptr, e->d_text_flags.ptr + (2 * buf + kind) * kMaxTextFlags,
                                                                         e->d_text_meta.ptr + 2 * kind + 1, kMaxTextFlags);
                CU_TRY(cudaGetLastError());
                ++res->launches;
            }
            CU_TRY(cudaMemcpyAsync(e->h_text_meta[buf].ptr, e->d_text_meta.ptr, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->compute));
            CU_TRY(cudaEventRecord(e->ev_text[buf], e->compute));
            // the number of records to re-format is only known now; the batch before this one is still crossing
            // PCIe / being written meanwhile
            CU_TRY(cudaEventSynchronize(e->ev_text[buf]));
            float tms = 0.f;
            CU_TRY(cudaEventElapsedTime(&tms, e->ev_computed[buf], e->ev_text[buf]));
            e->text_kernel_ms += tms;
            CU_TRY(cudaEventRecord(e->ev_copy_begin[buf], e->copy));
            for (int kind = 0; kind < 2; ++kind) {
                if (!n_text_slots[kind]) continue;
                const unsigned long long total = pending[buf].text_bytes[kind], nf = e->h_text_meta[buf].ptr[2 * kind + 1];
                if (nf > kMaxTextFlags) return fail(CPPROB_SIS_EIO, "device text stage: too many ambiguous records in one batch");
                CU_TRY(cudaMemcpyAsync(e->h_text[buf][kind].ptr, e->d_text[buf][kind].ptr, total, cudaMemcpyDeviceToHost, e->copy));
                if (nf) {
                    CU_TRY(cudaMemcpyAsync(e->h_text_flags.ptr + (2 * buf + kind) * kMaxTextFlags, e->d_text_flags.ptr + (2 * buf + kind) * kMaxTextFlags,
                                           nf * sizeof(text_flag), cudaMemcpyDeviceToHost, e->copy));
                }
                pending[buf].text_flags[kind] = nf;
            }
            CU_TRY(cudaEventRecord(e->ev_copied[buf], e->copy));
            pending[buf].valid = true;
            pending[buf].first = plan.first_particle + off;
            pending[buf].n = n_here;
            if (int rc = deliver(buf ^ 1)) return rc;
        } else if (emit) {
            CU_TRY(cudaEventRecord(e->ev_computed[buf], e->compute));
            CU_TRY(cudaStreamWaitEvent(e->copy, e->ev_computed[buf], 0));
            if (n_real > 0) {
                CU_TRY(cudaMemcpy2DAsync(e->h_real[buf].ptr, cap * sizeof(double), a.real_rows, cap * sizeof(double),
                                         n_here * sizeof(double), n_real, cudaMemcpyDeviceToHost, e->copy));
            }
            if (n_int > 0) {
                CU_TRY(cudaMemcpy2DAsync(e->h_int[buf].ptr, cap * sizeof(int), a.int_rows, cap * sizeof(int),
                                         n_here * sizeof(int), n_int, cudaMemcpyDeviceToHost, e->copy));
            }
            CU_TRY(cudaMemcpyAsync(e->h_logw[buf].ptr, a.logw, n_here * sizeof(double), cudaMemcpyDeviceToHost, e->copy));
            CU_TRY(cudaEventRecord(e->ev_copied[buf], e->copy));
            pending[buf].valid = true;
            pending[buf].first = plan.first_particle + off;
            pending[buf].n = n_here;
            // while this batch computes and copies, hand the previous one to the consumer
            if (int rc = deliver(buf ^ 1)) return rc;
        }
    }
    if (emit) {
        const int last = static_cast<int>((n_batches - 1) & 1);
        if (int rc = deliver(last ^ 1)) return rc;
        if (int rc = deliver(last)) return rc;
        if (int rc = fold_rows()) return rc;
        CU_TRY(cudaStreamSynchronize(e->compute));
    } else {
        if (int rc = fold_rows()) return rc;
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        res->device_ms = ms;
    }
    return 0;
}

constexpr double kRebaseLimit = 300.0;

// Merge [n_chunks][n_cols] partials (device) and turn the sums into estimators.
// Returns 1 if the weights must be re-based to out->max_log_w.
// `pending`: a fused shard of this engine that was launched without waiting (shard_options::no_wait); its m_ref and
// device time are collected here, after the one synchronisation of the inference.
int merge_impl(cpprob_sis_engine * e, const double * gathered, uint32_t n_chunks, int n_cols, int n_real, int n_int,
               hist_window hw, double m_ref, uint64_t n_total, cpprob_sis_stats * out, uint64_t * launches, double * ms_out,
               shard_result * pending = nullptr, const gather_layout * layout = nullptr,
               const unsigned long long * peer_flags = nullptr, unsigned long long peer_epoch = 0)
{
    if (n_cols != kBaseCols + 2 * n_real + n_int * hw.bins) return fail(CPPROB_SIS_EINVAL, "n_cols does not match the model structure");
    // The merge kernels store their n_cols (+ 2) results straight into pinned host memory (cudaMallocHost memory is mapped
    // into the device's address space under unified addressing): no copy-engine operation behind the last kernel.
    CU_TRY(e->h_merged.reserve(static_cast<size_t>(n_cols) + 2));
    CU_TRY(cudaEventRecord(e->ev_merge_begin, e->compute));
    // a particle pass that is still in flight left m_ref on the device (e->d_pilot[0]): it comes back behind the sums
    const bool m_ref_pending = pending && pending->waiting;
    const double * m_ref_dev = m_ref_pending ? e->d_pilot.ptr : nullptr;
    if (layout) {        // the raw output of an all-gather: read in place (k_merge_columns_gathered)
        k_merge_columns_gathered<<<n_cols, kBlock, 0, e->compute>>>(gathered, *layout, n_cols, kMaxColsMask, e->h_merged.ptr, m_ref_dev,
                                                                   peer_flags, peer_epoch);
    } else {
        k_merge_columns<<<n_cols, kBlock, 0, e->compute>>>(gathered, n_chunks, n_cols, kMaxColsMask, e->h_merged.ptr, m_ref_dev);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(e->ev_merge_end, e->compute));
    ++*launches;
    CU_TRY(cudaStreamSynchronize(e->compute));
    if (peer_flags && e->h_merged.ptr[n_cols + 1] != 0.0) {
        return fail(CPPROB_SIS_ENCCL, "a peer's partial rows did not arrive within 20 s (peer-memory exchange): a rank failed or left the collective");
    }
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e->ev_merge_begin, e->ev_merge_end));
    *ms_out = ms;
    if (m_ref_pending) {
        pending->waiting = false;
        pending->m_ref = m_ref = e->h_merged.ptr[n_cols];
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        pending->device_ms = ms;
    }

    const double * s = e->h_merged.ptr;
    e->sums.assign(s, s + n_cols);
    std::memset(out, 0, sizeof *out);
    out->n_particles = n_total;
    out->n_neg_inf = static_cast<uint64_t>(s[col::n_neginf]);
    out->n_nan = static_cast<uint64_t>(s[col::n_nan]);
    out->m_ref = m_ref;
    out->max_log_w = s[col::max_lw];
    // a NaN or +inf log-weight makes every self-normalised estimator NaN, as it does in the reference
    // (exp(nan - lse) / exp(inf - inf) in empirical_distribution.hpp:52-66); exp_weight does not
    // propagate those itself, so poison the sums here
    const bool poisoned = s[col::n_nan] > 0.0 || s[col::max_lw] == std::numeric_limits<double>::infinity();
    const double s0 = poisoned ? std::numeric_limits<double>::quiet_NaN() : s[col::s0], s00 = s[col::s00];
    out->log_sum_exp = m_ref + std::log(s0);
    out->log_evidence = out->log_sum_exp - std::log(static_cast<double>(n_total));
    out->ess = s0 * s0 / s00;
    out->n_real = n_real;
    out->n_int = n_int;
    e->real_mean.assign(static_cast<size_t>(n_real), 0.0);
    e->real_var.assign(static_cast<size_t>(n_real), 0.0);
    for (int j = 0; j < n_real; ++j) {
        // empirical_distribution.hpp:52-81: mean = sum w~ x, variance = sum w~ x^2 - mean*mean
        const double mean = s[kBaseCols + 2 * j] / s0;
        e->real_mean[static_cast<size_t>(j)] = mean;
        e->real_var[static_cast<size_t>(j)] = s[kBaseCols + 2 * j + 1] / s0 - mean * mean;
    }
    e->int_prob.assign(static_cast<size_t>(n_int) * hw.bins, 0.0);
    e->int_map.assign(static_cast<size_t>(n_int), 0);
    const int hist0 = kBaseCols + 2 * n_real;
    for (int j = 0; j < n_int; ++j) {
        int best = 0;
        for (int b = 0; b < hw.bins; ++b) {
            const double p = s[hist0 + j * hw.bins + b] / s0;
            e->int_prob[static_cast<size_t>(j) * hw.bins + b] = p;
            if (p > e->int_prob[static_cast<size_t>(j) * hw.bins + best]) best = b;   // first maximum wins
        }
        e->int_map[static_cast<size_t>(j)] = hw.lo + best;
    }
    out->real_mean = e->real_mean.data();
    out->real_var = e->real_var.data();
    out->int_lo = hw.lo;
    out->int_bins = hw.bins;
    out->int_prob = e->int_prob.data();
    out->int_map = e->int_map.data();
    out->n_cols = n_cols;
    out->sums = e->sums.data();

    // re-base if exp(log_w - m_ref) could have overflowed / underflowed the interesting weights.  The tightest
    // quantity is sum w^2 (the ESS denominator): w^2 overflows a double once log_w - m_ref exceeds 354.9, so the
    // limit is 300 (weights up to e^300, squares up to e^600, 2^43 of them still finite)
    const double mx = out->max_log_w;
    if (std::isfinite(mx) && std::fabs(mx - m_ref) > kRebaseLimit) return 1;
    return 0;
}

// window that covers every int value seen (from merged columns); false if it already did
bool widen_window(const cpprob_sis_stats & st, int n_int, hist_window * hw)
{
    if (n_int == 0) return false;
    if (st.sums[col::int_oor] == 0.0) return false;
    const double lo = -st.sums[col::neg_imin], hi = st.sums[col::imax];
    hw->lo = static_cast<long long>(lo);
    hw->bins = static_cast<int>(std::min<double>(hi - lo + 1.0, 1.0e9));
    return true;
}

int run_full(cpprob_sis_engine * e, const cpprob_sis_model_vtable * vt, const double * obs, size_t n_obs, uint64_t n,
             const shard_options & opt_in, cpprob_sis_stats * out)
{
    shard_options opt = opt_in;
    shard_result res;
    double m_ref_override = 0.0;
    const double * mo = nullptr;
    hist_window hw_override;
    const hist_window * ho = nullptr;
    double total_ms = 0.0, particle_ms = 0.0;
    uint64_t launches = 0;
    int passes = 0;
    opt.no_wait = true;      // one synchronisation per inference on the fused path: after the merge
    for (;;) {
        res.waiting = false;
        if (int rc = run_shard_impl(e, vt, obs, n_obs, n, 0, 1, mo, ho, opt, &res)) return rc;
        ++passes;
        launches += res.launches;
        double merge_ms = 0.0;
        const int rc = merge_impl(e, res.rows, res.n_rows_total, res.n_cols, static_cast<int>(e->structure.n_real),
                                  static_cast<int>(e->structure.n_int), res.hw, res.m_ref, n, out, &launches, &merge_ms, &res);
        if (rc < 0) return rc;
        total_ms += res.device_ms + merge_ms;
        particle_ms += res.device_ms;
        bool again = false;
        if (rc == 1 && passes < 3) {
            m_ref_override = out->max_log_w;
            mo = &m_ref_override;
            again = true;
        }
        hist_window hw = res.hw;
        if (vt->int_states > 0 && e->structure.n_int > 0 && out->sums[col::int_oor] != 0.0) {
            return fail(CPPROB_SIS_ERANGE, std::string("model ") + vt->name + " predicted an integral value outside [0, " + std::to_string(vt->int_states) +
                                               "), the range it declares (int_predict_states)");
        }
        if (passes < 3 && widen_window(*out, static_cast<int>(e->structure.n_int), &hw)) {
            if (hw.bins > 4096) return fail(CPPROB_SIS_ERANGE, "int predicts span more than 4096 values");
            hw_override = hw;
            ho = &hw_override;
            again = true;
        }
        if (!again) break;
        // the records were already delivered in the first pass; later passes only redo the sums
        opt.emit = CPPROB_SIS_EMIT_NONE;
        opt.on_block = nullptr;
        opt.text_writer = nullptr;
        opt.force_rows = opt_in.force_rows || opt_in.emit == CPPROB_SIS_EMIT_ALL;
    }
    out->device_ms = total_ms;
    out->particle_ms = particle_ms;
    out->kernel_launches = launches;
    out->passes = passes;
    out->path = res.path;
    e->launches += launches;
    return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

static int merge_gathered(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * gathered,
                          const gather_layout * lay, int n_cols, double m_ref, uint64_t n_particles_total, cpprob_sis_stats * out);

int cpprob_sis_abi_version(void) { return CPPROB_SIS_ABI_VERSION; }

const char * cpprob_sis_last_error(void) { return g_last_error.c_str(); }

int cpprob_sis_register_model(const cpprob_sis_model_vtable * vt)
{
    if (!vt || vt->abi_version != CPPROB_SIS_ABI_VERSION || !vt->name) return fail(CPPROB_SIS_EINVAL, "bad model vtable");
    std::lock_guard<std::mutex> lock(registry_mutex());
    for (size_t i = 0; i < registry().size(); ++i) {
        if (std::strcmp(registry()[i]->name, vt->name) == 0) {
            registry()[i] = vt;
            return static_cast<int>(i);
        }
    }
    registry().push_back(vt);
    return static_cast<int>(registry().size() - 1);
}

int cpprob_sis_model_count(void)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    return static_cast<int>(registry().size());
}

const char * cpprob_sis_model_name(int model_id)
{
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    return vt ? vt->name : nullptr;
}

int cpprob_sis_find_model(const char * name)
{
    if (!name) return fail(CPPROB_SIS_EINVAL, "null model name");
    std::lock_guard<std::mutex> lock(registry_mutex());
    for (size_t i = 0; i < registry().size(); ++i) {
        if (std::strcmp(registry()[i]->name, name) == 0) return static_cast<int>(i);
    }
    return fail(CPPROB_SIS_ENOMODEL, std::string("no device model named '") + name + "' is registered");
}

int cpprob_sis_plan_shard(uint64_t n_particles_total, int rank, int world, uint32_t * chunk_first, uint32_t * n_chunks_local,
                          uint32_t * n_chunks_total, uint64_t * first_particle, uint64_t * n_local)
{
    if (world <= 0 || rank < 0 || rank >= world || n_particles_total == 0) return fail(CPPROB_SIS_EINVAL, "bad rank / world / n");
    const shard_plan p = plan_shard(n_particles_total, rank, world);
    if (chunk_first) *chunk_first = p.chunk_first;
    if (n_chunks_local) *n_chunks_local = p.n_chunks_local;
    if (n_chunks_total) *n_chunks_total = p.n_chunks_total;
    if (first_particle) *first_particle = p.first_particle;
    if (n_local) *n_local = p.n_local;
    return 0;
}

int cpprob_sis_plan_rows(uint64_t n_particles_total, int rank, int world, int rows_per_chunk, uint32_t * row_first,
                         uint32_t * n_rows_local, uint32_t * n_rows_total)
{
    if (world <= 0 || rank < 0 || rank >= world || n_particles_total == 0 || rows_per_chunk <= 0) {
        return fail(CPPROB_SIS_EINVAL, "bad rank / world / n / rows_per_chunk");
    }
    const shard_plan p = plan_shard(n_particles_total, rank, world);
    if (p.super * static_cast<uint32_t>(rows_per_chunk) == 1) {
        if (row_first) *row_first = p.chunk_first;
        if (n_rows_local) *n_rows_local = p.n_chunks_local;
        if (n_rows_total) *n_rows_total = p.n_chunks_total;
    } else {
        if (row_first) *row_first = p.super_first;
        if (n_rows_local) *n_rows_local = p.n_super_local;
        if (n_rows_total) *n_rows_total = p.n_super_total;
    }
    return 0;
}

int cpprob_sis_create(const cpprob_sis_config * cfg, cpprob_sis_engine ** out)
{
    if (!out) return fail(CPPROB_SIS_EINVAL, "null out pointer");
    *out = nullptr;
    int n_dev = 0;
    const cudaError_t err = cudaGetDeviceCount(&n_dev);
    if (err != cudaSuccess || n_dev == 0) {
        return fail(CPPROB_SIS_ECUDA, std::string("no usable CUDA device (the SIS engine has no CPU fallback): ") +
                                          (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0"));
    }
    cpprob_sis_engine * e = new cpprob_sis_engine();
    e->device = cfg ? cfg->device : 0;
    e->seed = cfg ? cfg->seed : 0x5eedull;
    e->max_batch = cfg ? cfg->max_batch / kChunk * kChunk : 0;
    e->blocks_per_sm = cfg ? cfg->blocks_per_sm : 0;
    if (e->device < 0 || e->device >= n_dev) {
        delete e;
        return fail(CPPROB_SIS_EINVAL, "device ordinal out of range");
    }
    auto bail = [&](cudaError_t cerr, const char * what) {
        const std::string msg = std::string(what) + ": " + cudaGetErrorString(cerr);
        cpprob_sis_destroy(e);
        return fail(CPPROB_SIS_ECUDA, msg);
    };
    cudaError_t c;
    if ((c = cudaSetDevice(e->device)) != cudaSuccess) return bail(c, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((c = cudaGetDeviceProperties(&prop, e->device)) != cudaSuccess) return bail(c, "cudaGetDeviceProperties");
    e->sm_count = prop.multiProcessorCount;
    if ((c = cudaStreamCreateWithFlags(&e->compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
    if ((c = cudaStreamCreateWithFlags(&e->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
    if ((c = cudaEventCreate(&e->ev_begin)) != cudaSuccess) return bail(c, "cudaEventCreate");
    if ((c = cudaEventCreate(&e->ev_end)) != cudaSuccess) return bail(c, "cudaEventCreate");
    if ((c = cudaEventCreate(&e->ev_merge_begin)) != cudaSuccess) return bail(c, "cudaEventCreate");
    if ((c = cudaEventCreate(&e->ev_merge_end)) != cudaSuccess) return bail(c, "cudaEventCreate");
    for (int i = 0; i < 2; ++i) {
        if ((c = cudaEventCreate(&e->ev_computed[i])) != cudaSuccess) return bail(c, "cudaEventCreate");
        if ((c = cudaEventCreate(&e->ev_copied[i])) != cudaSuccess) return bail(c, "cudaEventCreate");
        if ((c = cudaEventCreate(&e->ev_batch_begin[i])) != cudaSuccess) return bail(c, "cudaEventCreate");
        if ((c = cudaEventCreate(&e->ev_text[i])) != cudaSuccess) return bail(c, "cudaEventCreate");
        if ((c = cudaEventCreate(&e->ev_copy_begin[i])) != cudaSuccess) return bail(c, "cudaEventCreate");
    }
    if (const char * s = std::getenv("CPPROB_SIS_TEXT_FORCE_AMBIGUOUS")) e->text_force_every = std::strtoull(s, nullptr, 10);
    *out = e;
    return 0;
}

int cpprob_sis_set_seed(cpprob_sis_engine * e, uint64_t seed)
{
    if (!e) return fail(CPPROB_SIS_EINVAL, "null engine");
    e->seed = seed;
    return 0;
}

void cpprob_sis_destroy(cpprob_sis_engine * e)
{
    if (!e) return;
    cpprob_sis_comm_destroy(e);
    cudaSetDevice(e->device);
    if (e->compute) cudaStreamSynchronize(e->compute);
    if (e->copy) cudaStreamSynchronize(e->copy);
    e->d_obs.release(); e->d_pilot.release(); e->d_partials.release(); e->d_super.release(); e->d_warp_partials.release(); e->d_gather.release();
    e->d_counter.release(); e->d_int_extra.release(); e->h_merged.release(); e->h_pilot.release(); e->h_obs.release();
    e->d_text_meta.release(); e->d_text_flags.release(); e->h_text_flags.release();
    for (int i = 0; i < 2; ++i) {
        e->d_text_slots[i].release(); e->h_text_meta[i].release(); e->d_text_len[i].release(); e->d_text_bsum[i].release();
        for (int k = 0; k < 2; ++k) { e->d_text[i][k].release(); e->h_text[i][k].release(); }
        if (e->ev_text[i]) cudaEventDestroy(e->ev_text[i]);
        if (e->ev_copy_begin[i]) cudaEventDestroy(e->ev_copy_begin[i]);
    }
    for (int i = 0; i < 2; ++i) {
        e->d_w[i].release(); e->d_logw[i].release(); e->d_real[i].release(); e->d_int[i].release();
        e->h_real[i].release(); e->h_logw[i].release(); e->h_int[i].release();
        if (e->ev_computed[i]) cudaEventDestroy(e->ev_computed[i]);
        if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
        if (e->ev_batch_begin[i]) cudaEventDestroy(e->ev_batch_begin[i]);
    }
    if (e->ev_merge_begin) cudaEventDestroy(e->ev_merge_begin);
    if (e->ev_merge_end) cudaEventDestroy(e->ev_merge_end);
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->compute) cudaStreamDestroy(e->compute);
    if (e->copy) cudaStreamDestroy(e->copy);
    delete e;
}

int cpprob_sis_describe(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, cpprob_sis_structure * out)
{
    if (!e || !out || !obs) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    if (int rc = probe_structure(e, vt, obs, n_obs)) return rc;
    out->n_ids = static_cast<int>(e->structure.ids.size());
    out->n_slots = static_cast<int>(e->slots.size());
    out->n_real = static_cast<int>(e->structure.n_real);
    out->n_int = static_cast<int>(e->structure.n_int);
    out->n_samples = static_cast<int>(e->structure.n_samples);
    out->ids = e->id_ptrs.data();
    out->slots = e->slots.data();
    return 0;
}

int cpprob_sis_run(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, uint64_t n_particles,
                   const cpprob_sis_run_options * opt, cpprob_sis_stats * out)
{
    if (!e || !out || !obs) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    shard_options so;
    if (opt) {
        so.emit = opt->emit;
        so.force_rows = opt->force_rows;
        so.on_block = opt->on_block;
        so.user = opt->user;
    }
    return run_full(e, vt, obs, n_obs, n_particles, so, out);
}

int cpprob_sis_run_shard(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, uint64_t n_particles_total,
                         int rank, int world, const double * m_ref_override, cpprob_sis_partials * out)
{
    if (!e || !out || !obs) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    shard_result res;
    shard_options so;
    if (int rc = run_shard_impl(e, vt, obs, n_obs, n_particles_total, rank, world, m_ref_override, nullptr, so, &res)) return rc;
    out->device_ptr = const_cast<double *>(res.rows);
    out->n_chunks_local = res.n_rows_local;
    out->n_chunks_total = res.n_rows_total;
    out->chunk_first = res.row_first;
    out->rows_per_chunk = res.rows_per_chunk;
    out->n_cols = res.n_cols;
    out->m_ref = res.m_ref;
    out->device_ms = res.device_ms;
    out->kernel_launches = res.launches;
    e->launches += res.launches;
    return 0;
}

namespace {
// An estimator-only run leaves no record file, and must not leave an older run's either: StatsPrinter reads the record
// files first and the .stats sidecar only when there are none, so stale records would be printed against the new .ids.
// (The reference's finish_infer removes the kinds that stayed empty, src/cpprob/state.cpp:167-175; here all of them did.)
void remove_record_files(const char * prefix)
{
    for (const char * ext : {".real", ".int", ".any"}) std::remove((std::string(prefix) + ext).c_str());
}
}  // namespace

int cpprob_sis_write_summary(cpprob_sis_engine * e, const char * prefix, const cpprob_sis_stats * stats)
{
    if (!e || !prefix || !stats) return fail(CPPROB_SIS_EINVAL, "null argument");
    remove_record_files(prefix);
    if (!cpprob::text::write_ids(prefix, e->structure.ids)) return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".ids");
    if (!cpprob::text::write_stats_sidecar(prefix, *stats, e->slots, e->structure.ids)) {
        return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".stats");
    }
    return 0;
}

// ---- multi-GPU inside the library ------------------------------------------------------------------------------------
namespace {

// One inference over the ranks of a communicator.  `local`: the engines of THIS process (one per GPU; one in the usual
// process-per-GPU set-up), each carrying its rank of the same communicator.  Every rank queues its shard, the ONE
// collective of the path — an all-gather of the per-(super-)chunk partial rows, straight from the buffer the shard
// kernels wrote — and the merge on its own stream; the host synchronises once, on local[0], whose engine owns the results.
// Re-basing / a wider histogram window repeat the pass on every rank alike (the decision is a function of the merged
// sums, which are bit-identical everywhere).
int run_dist_impl(cpprob_sis_engine * const * local, int n_local, const cpprob_sis_model_vtable * vt, const double * obs, size_t n_obs,
                  uint64_t n_total, cpprob_sis_stats * out)
{
    const nccl_api & nc = nccl();
    const int world = local[0]->comm_world;
    if (world > 1 && local[0]->comm && !nc.handle) return fail(CPPROB_SIS_ENCCL, nc.why);
    for (int i = 0; i < n_local; ++i) {
        if (world > 1 && !local[i]->comm && !local[i]->pw.ready) return fail(CPPROB_SIS_EINVAL, "engine has no communicator (cpprob_sis_comm_init)");
        if (local[i]->comm_world != world) return fail(CPPROB_SIS_EINVAL, "engines belong to different communicators");
        if (local[i]->seed != local[0]->seed) return fail(CPPROB_SIS_EINVAL, "all engines of a multi-GPU run must share one seed");
    }
    cpprob_sis_engine * primary = local[0];
    double m_ref_override = 0.0;
    const double * mo = nullptr;
    hist_window hw_override;
    const hist_window * ho = nullptr;
    double total_ms = 0.0, particle_ms = 0.0;
    uint64_t launches = 0;
    std::vector<shard_result> res(static_cast<size_t>(n_local));
    for (int pass = 1;; ++pass) {
        shard_options so;
        so.no_wait = true;
        so.push_peers = true;
        for (int i = 0; i < n_local; ++i) {
            res[static_cast<size_t>(i)] = shard_result();
            if (int rc = run_shard_impl(local[i], vt, obs, n_obs, n_total, local[i]->comm_rank, world, mo, ho, so, &res[static_cast<size_t>(i)])) return rc;
            launches += res[static_cast<size_t>(i)].launches;
        }
        const shard_result & r0 = res[0];
        const int n_cols = r0.n_cols;
        gather_layout lay;
        if (int rc = make_gather_layout(n_total, world, static_cast<int>(r0.rows_per_chunk), &lay)) return rc;
        const size_t count = static_cast<size_t>(lay.rows_per_rank) * n_cols;
        const double * merged_from = nullptr;
        // the exchange: through the ranks' peer windows (the shard's last kernel stored the rows there and raised the flags;
        // the merge kernel waits for them) when every rank has them and the rows fit, else one ncclAllGather.  Both
        // decisions come out the same on every rank.
        bool use_peer = world > 1;
        for (int i = 0; i < n_local; ++i) use_peer = use_peer && res[static_cast<size_t>(i)].pushed;
        const unsigned long long * peer_flags = nullptr;
        unsigned long long peer_epoch = 0;
        if (use_peer) {                            // the shard kernels pushed the rows themselves
            merged_from = reinterpret_cast<const double *>(primary->pw.local + r0.push_offset);
            peer_flags = reinterpret_cast<const unsigned long long *>(primary->pw.local);
            peer_epoch = r0.push_epoch;
        } else if (world > 1) {
            if (!primary->comm) return fail(CPPROB_SIS_EINVAL, "the partial rows of this run do not fit the peer windows and the engines have no NCCL communicator");
            if (n_local > 1) NCCL_TRY(nc.GroupStart());
            for (int i = 0; i < n_local; ++i) {
                cpprob_sis_engine * e = local[i];
                if (int rc = use_device(e)) return rc;
                CU_TRY(e->d_gather.reserve(count * static_cast<size_t>(world)));
                const double * send = res[static_cast<size_t>(i)].rows;
                if (!send) {                       // a rank without particles still takes part; its segment is never read
                    CU_TRY(e->d_partials.reserve(count + static_cast<size_t>(n_cols)));
                    send = e->d_partials.ptr;
                }
                NCCL_TRY(nc.AllGather(send, e->d_gather.ptr, count, ncclDouble, e->comm, e->compute));
            }
            if (n_local > 1) NCCL_TRY(nc.GroupEnd());
            merged_from = primary->d_gather.ptr;
        }
        if (int rc = use_device(primary)) return rc;
        double merge_ms = 0.0;
        int mrc;
        if (world > 1) {
            mrc = merge_impl(primary, merged_from, lay.first[world], n_cols, static_cast<int>(primary->structure.n_real),
                             static_cast<int>(primary->structure.n_int), r0.hw, r0.m_ref, n_total, out, &launches, &merge_ms, &res[0], &lay,
                             peer_flags, peer_epoch);
        } else {
            mrc = merge_impl(primary, r0.rows, r0.n_rows_total, n_cols, static_cast<int>(primary->structure.n_real),
                             static_cast<int>(primary->structure.n_int), r0.hw, r0.m_ref, n_total, out, &launches, &merge_ms, &res[0]);
        }
        if (mrc < 0) return mrc;
        // the other local engines only have to finish (their device time is taken after the fact; max over ranks)
        double shard_ms = res[0].device_ms;
        for (int i = 1; i < n_local; ++i) {
            cpprob_sis_engine * e = local[i];
            if (int rc = use_device(e)) return rc;
            CU_TRY(cudaStreamSynchronize(e->compute));
            if (res[static_cast<size_t>(i)].waiting) {
                float ms = 0.f;
                CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
                res[static_cast<size_t>(i)].device_ms = ms;
                res[static_cast<size_t>(i)].waiting = false;
            }
            shard_ms = std::max(shard_ms, res[static_cast<size_t>(i)].device_ms);
        }
        total_ms += shard_ms + merge_ms;
        particle_ms += res[0].device_ms;
        bool again = false;
        if (mrc == 1 && pass < 3) {
            m_ref_override = out->max_log_w;
            mo = &m_ref_override;
            again = true;
        }
        hist_window hw = r0.hw;
        if (vt->int_states > 0 && primary->structure.n_int > 0 && out->sums[col::int_oor] != 0.0) {
            return fail(CPPROB_SIS_ERANGE, std::string("model ") + vt->name + " predicted an integral value outside [0, " + std::to_string(vt->int_states) +
                                               "), the range it declares (int_predict_states)");
        }
        if (pass < 3 && widen_window(*out, static_cast<int>(primary->structure.n_int), &hw)) {
            if (hw.bins > 4096) return fail(CPPROB_SIS_ERANGE, "int predicts span more than 4096 values");
            hw_override = hw;
            ho = &hw_override;
            again = true;
        }
        if (!again) {
            if (primary->structure.n_int > 0 && out->sums[col::int_oor] != 0.0) {
                return fail(CPPROB_SIS_ERANGE, "int predicts fell outside the histogram window");
            }
            out->passes = pass;
            out->path = r0.path;
            break;
        }
    }
    out->device_ms = total_ms;
    out->particle_ms = particle_ms;
    out->kernel_launches = launches;
    primary->launches += launches;
    return 0;
}

}  // namespace

namespace {

constexpr size_t kPeerBufferBytes = 8u << 20;      // one gather buffer of a peer window (two per window)

bool peer_exchange_wanted()
{
    const char * v = std::getenv("CPPROB_SIS_EXCHANGE");       // "nccl": always ncclAllGather; default: peer memory where possible
    return !(v && std::strcmp(v, "nccl") == 0);
}

void peer_window_release(cpprob_sis_engine * e)
{
    cpprob_sis_engine::peer_window & w = e->pw;
    for (int p = 0; p < kMaxMergeRanks; ++p) {
        if (w.opened[p] && w.peer[p]) cudaIpcCloseMemHandle(w.peer[p]);
        w.opened[p] = false;
        w.peer[p] = nullptr;
    }
    if (w.local) cudaFree(w.local);
    w = cpprob_sis_engine::peer_window();
}

cudaError_t peer_window_alloc(cpprob_sis_engine * e)
{
    cpprob_sis_engine::peer_window & w = e->pw;
    w.buffer_bytes = kPeerBufferBytes;
    const size_t bytes = kPeerFlagBytes + 2 * w.buffer_bytes;
    cudaError_t err = cudaMalloc(reinterpret_cast<void **>(&w.local), bytes);
    if (err != cudaSuccess) { w.local = nullptr; return err; }
    err = cudaMemset(w.local, 0, bytes);
    return err;
}

// One process per GPU: the windows are exchanged as CUDA IPC handles over the communicator that was just made, and the
// peer path is switched on only if EVERY rank mapped EVERY window (the ranks must agree, or one would wait in NCCL for
// peers that push through memory).  Any failure leaves the NCCL path in place; it is not an error.
int peer_window_setup_ipc(cpprob_sis_engine * e)
{
    const nccl_api & nc = nccl();
    const int world = e->comm_world, rank = e->comm_rank;
    struct record { cudaIpcMemHandle_t handle; unsigned long long ok; };
    static_assert(sizeof(record) == 72, "IPC handle record");
    record mine;
    std::memset(&mine, 0, sizeof mine);
    bool ok = peer_exchange_wanted() && peer_window_alloc(e) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, e->pw.local) == cudaSuccess;
    cudaGetLastError();
    mine.ok = ok ? 1ull : 0ull;
    device_buffer<unsigned char> d_rec;
    CU_TRY(d_rec.reserve(sizeof(record) * static_cast<size_t>(world)));
    std::vector<record> all(static_cast<size_t>(world));
    auto gather = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(d_rec.ptr + sizeof(record) * static_cast<size_t>(rank), &mine, sizeof mine, cudaMemcpyHostToDevice, e->compute));
        NCCL_TRY(nc.AllGather(d_rec.ptr + sizeof(record) * static_cast<size_t>(rank), d_rec.ptr, sizeof(record), ncclChar, e->comm, e->compute));
        CU_TRY(cudaMemcpyAsync(all.data(), d_rec.ptr, sizeof(record) * static_cast<size_t>(world), cudaMemcpyDeviceToHost, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        return 0;
    };
    int rc = gather();
    if (rc) { d_rec.release(); return rc; }
    bool everyone = true;
    for (const record & r : all) everyone = everyone && r.ok == 1ull;
    if (everyone) {
        for (int p = 0; p < world && ok; ++p) {
            if (p == rank) { e->pw.peer[p] = e->pw.local; continue; }
            void * ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[static_cast<size_t>(p)].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                break;
            }
            e->pw.peer[p] = static_cast<unsigned char *>(ptr);
            e->pw.opened[p] = true;
        }
    }
    // second round: did every rank map every window?
    mine.ok = (everyone && ok) ? 1ull : 0ull;
    rc = gather();
    d_rec.release();
    if (rc) return rc;
    bool ready = true;
    for (const record & r : all) ready = ready && r.ok == 1ull;
    if (ready) e->pw.ready = true;
    else peer_window_release(e);
    return 0;
}

// One process, one engine per GPU: peer access between the devices, plain pointers.
void peer_window_setup_local(cpprob_sis_engine * const * engines, int n)
{
    if (!peer_exchange_wanted()) return;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        if (cudaSetDevice(engines[i]->device) != cudaSuccess) { ok = false; break; }
        for (int j = 0; j < n && ok; ++j) {
            if (i == j || engines[i]->device == engines[j]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, engines[i]->device, engines[j]->device) != cudaSuccess || !can) { ok = false; break; }
            const cudaError_t err = cudaDeviceEnablePeerAccess(engines[j]->device, 0);
            if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
        if (ok) ok = peer_window_alloc(engines[i]) == cudaSuccess;
    }
    cudaGetLastError();
    if (!ok) {
        for (int i = 0; i < n; ++i) { cudaSetDevice(engines[i]->device); peer_window_release(engines[i]); }
        return;
    }
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) engines[i]->pw.peer[j] = engines[j]->pw.local;
        engines[i]->pw.ready = true;
    }
}

}  // namespace

int cpprob_sis_comm_get_id(void * id_out)
{
    if (!id_out) return fail(CPPROB_SIS_EINVAL, "null argument");
    const nccl_api & nc = nccl();
    if (!nc.handle) return fail(CPPROB_SIS_ENCCL, nc.why);
    static_assert(sizeof(ncclUniqueId) == CPPROB_SIS_COMM_ID_BYTES, "cpprob_sis.h states the size of an NCCL unique id");
    ncclUniqueId id;
    NCCL_TRY(nc.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return 0;
}

int cpprob_sis_comm_init(cpprob_sis_engine * e, const void * id_in, int rank, int world)
{
    if (!e || !id_in || world <= 0 || rank < 0 || rank >= world) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (world > kMaxMergeRanks) return fail(CPPROB_SIS_EINVAL, "more than 64 ranks");
    const nccl_api & nc = nccl();
    if (!nc.handle) return fail(CPPROB_SIS_ENCCL, nc.why);
    if (int rc = cpprob_sis_comm_destroy(e)) return rc;
    if (int rc = use_device(e)) return rc;
    ncclUniqueId id;
    std::memcpy(&id, id_in, sizeof id);
    NCCL_TRY(nc.CommInitRank(&e->comm, world, id, rank));
    e->comm_rank = rank;
    e->comm_world = world;
    if (world > 1) {
        if (int rc = peer_window_setup_ipc(e)) return rc;
    }
    return 0;
}

int cpprob_sis_comm_init_local(cpprob_sis_engine * const * engines, int n_engines)
{
    if (!engines || n_engines <= 0 || n_engines > kMaxMergeRanks) return fail(CPPROB_SIS_EINVAL, "bad argument");
    std::vector<int> devs;
    bool shared_device = false;
    for (int r = 0; r < n_engines; ++r) {
        if (!engines[r]) return fail(CPPROB_SIS_EINVAL, "null engine");
        for (int d : devs) shared_device = shared_device || d == engines[r]->device;
        devs.push_back(engines[r]->device);
        if (int rc = cpprob_sis_comm_destroy(engines[r])) return rc;
    }
    if (n_engines == 1) return 0;                       // rank 0 of 1 needs no communicator
    // Engines that share a device can be the ranks of one run only through the peer windows (NCCL refuses two ranks on
    // one device): several shards of a run on one GPU — what the single-GPU tests use to exercise the exchange.
    if (!shared_device) {
        const nccl_api & nc = nccl();
        if (!nc.handle) return fail(CPPROB_SIS_ENCCL, nc.why);
        std::vector<ncclComm_t> comms(static_cast<size_t>(n_engines));
        NCCL_TRY(nc.CommInitAll(comms.data(), n_engines, devs.data()));
        for (int r = 0; r < n_engines; ++r) engines[r]->comm = comms[static_cast<size_t>(r)];
    }
    for (int r = 0; r < n_engines; ++r) {
        engines[r]->comm_rank = r;
        engines[r]->comm_world = n_engines;
    }
    peer_window_setup_local(engines, n_engines);
    if (shared_device && !engines[0]->pw.ready) {
        for (int r = 0; r < n_engines; ++r) cpprob_sis_comm_destroy(engines[r]);
        return fail(CPPROB_SIS_EINVAL, "engines on one device can only share a communicator through peer windows (CPPROB_SIS_EXCHANGE=nccl or no memory for them)");
    }
    return 0;
}

int cpprob_sis_comm_destroy(cpprob_sis_engine * e)
{
    if (!e) return fail(CPPROB_SIS_EINVAL, "null engine");
    if (e->comm) {
        cudaSetDevice(e->device);
        if (e->compute) cudaStreamSynchronize(e->compute);
        nccl().CommDestroy(e->comm);
        e->comm = nullptr;
    }
    if (e->pw.local) {
        cudaSetDevice(e->device);
        if (e->compute) cudaStreamSynchronize(e->compute);
        peer_window_release(e);
    }
    e->comm_rank = 0;
    e->comm_world = 1;
    return 0;
}

int cpprob_sis_comm_exchange(const cpprob_sis_engine * e)
{
    if (!e || e->comm_world <= 1) return CPPROB_SIS_EXCHANGE_NONE;
    return e->pw.ready ? CPPROB_SIS_EXCHANGE_PEER : CPPROB_SIS_EXCHANGE_NCCL;
}

int cpprob_sis_run_dist(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, uint64_t n_particles_total, cpprob_sis_stats * out)
{
    if (!e || !obs || !out) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    return run_dist_impl(&e, 1, vt, obs, n_obs, n_particles_total, out);
}

int cpprob_sis_run_multi(cpprob_sis_engine * const * engines, int n_engines, int model_id, const double * obs, size_t n_obs,
                         uint64_t n_particles, cpprob_sis_stats * out)
{
    if (!engines || n_engines <= 0 || !obs || !out) return fail(CPPROB_SIS_EINVAL, "bad argument");
    for (int r = 0; r < n_engines; ++r) if (!engines[r]) return fail(CPPROB_SIS_EINVAL, "null engine");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    // the engines need one communicator among themselves, rank r = position in the list; made on first use and kept
    bool ready = true;
    for (int r = 0; r < n_engines; ++r) {
        ready = ready && engines[r]->comm_world == n_engines && engines[r]->comm_rank == r && (n_engines == 1 || engines[r]->comm || engines[r]->pw.ready);
    }
    if (!ready) {
        if (int rc = cpprob_sis_comm_init_local(engines, n_engines)) return rc;
    }
    return run_dist_impl(engines, n_engines, vt, obs, n_obs, n_particles, out);
}

int cpprob_sis_merge_padded(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * gathered,
                            int world, uint32_t rows_per_rank, int rows_per_chunk, int n_cols, double m_ref,
                            uint64_t n_particles_total, cpprob_sis_stats * out)
{
    if (!e || !gathered || !obs || !out || world <= 0 || n_cols <= 0) return fail(CPPROB_SIS_EINVAL, "bad argument");
    gather_layout lay;
    if (int rc = make_gather_layout(n_particles_total, world, rows_per_chunk, &lay)) return rc;
    for (int r = 0; r < world; ++r) {
        if (lay.first[r + 1] - lay.first[r] > rows_per_rank) return fail(CPPROB_SIS_EINVAL, "rows_per_rank is smaller than a rank's row count");
    }
    lay.rows_per_rank = rows_per_rank;
    return merge_gathered(e, model_id, obs, n_obs, gathered, &lay, n_cols, m_ref, n_particles_total, out);
}

int cpprob_sis_merge(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * gathered,
                     uint32_t n_chunks_total, int n_cols, double m_ref, uint64_t n_particles_total, cpprob_sis_stats * out)
{
    gather_layout lay;                       // one segment holding every row
    lay.world = 1;
    lay.first[0] = 0;
    lay.first[1] = n_chunks_total;
    lay.rows_per_rank = n_chunks_total;
    return merge_gathered(e, model_id, obs, n_obs, gathered, &lay, n_cols, m_ref, n_particles_total, out);
}

static int merge_gathered(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * gathered,
                          const gather_layout * lay, int n_cols, double m_ref, uint64_t n_particles_total, cpprob_sis_stats * out)
{
    const uint32_t n_chunks_total = lay->first[lay->world];
    if (!e || !out || !obs || !gathered) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    if (int rc = use_device(e)) return rc;
    if (int rc = probe_structure(e, vt, obs, n_obs)) return rc;
    const int n_real = static_cast<int>(e->structure.n_real), n_int = static_cast<int>(e->structure.n_int);
    hist_window hw;
    if (n_int > 0) {
        // the window is a pure function of (seed, model, obs): recompute it the way run_shard did
        const int rest = n_cols - kBaseCols - 2 * n_real;
        if (rest <= 0 || rest % n_int != 0) return fail(CPPROB_SIS_EINVAL, "n_cols does not match the model structure");
        hw.bins = rest / n_int;
        CU_TRY(e->d_obs.reserve(n_obs));
        CU_TRY(e->d_pilot.reserve(4));
        CU_TRY(cudaMemcpyAsync(e->d_obs.ptr, obs, n_obs * sizeof(double), cudaMemcpyHostToDevice, e->compute));
        const philox_keys keys(e->seed);
        double pilot[3];
        if (int rc = run_pilot(e, vt, keys, n_obs, n_particles_total, nullptr, pilot)) return rc;
        hw.lo = pilot[1] <= pilot[2] ? static_cast<long long>(pilot[1]) : 0;
    }
    uint64_t launches = 0;
    double ms = 0.0;
    const int rc = merge_impl(e, gathered, n_chunks_total, n_cols, n_real, n_int, hw, m_ref, n_particles_total, out, &launches, &ms, nullptr,
                              lay->world > 1 ? lay : nullptr);
    if (rc < 0) return rc;
    if (n_int > 0 && out->sums[col::int_oor] != 0.0) {
        return fail(CPPROB_SIS_ERANGE, "int predicts fell outside the pilot's histogram window");
    }
    out->device_ms = ms;
    out->kernel_launches = launches;
    out->passes = 1;
    e->launches += launches;
    return rc;
}

int cpprob_sis_replay(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, const double * real_rows,
                      const int32_t * int_rows, uint64_t stride, uint64_t n, double * logw_out)
{
    if (!e || !obs || !logw_out) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    if (!vt->replayable) return fail(CPPROB_SIS_EINVAL, std::string("model ") + vt->name + " is not replayable");
    if (n == 0) return 0;
    if (int rc = use_device(e)) return rc;
    if (int rc = probe_structure(e, vt, obs, n_obs)) return rc;
    const size_t n_real = e->structure.n_real, n_int = e->structure.n_int;
    if ((n_real && !real_rows) || (n_int && !int_rows) || stride < n) return fail(CPPROB_SIS_EINVAL, "bad trace rows");
    CU_TRY(e->d_obs.reserve(n_obs));
    CU_TRY(e->d_real[0].reserve(std::max<size_t>(1, n_real * stride)));
    CU_TRY(e->d_int[0].reserve(std::max<size_t>(1, n_int * stride)));
    CU_TRY(e->d_logw[0].reserve(n));
    CU_TRY(cudaMemcpyAsync(e->d_obs.ptr, obs, n_obs * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    if (n_real) CU_TRY(cudaMemcpyAsync(e->d_real[0].ptr, real_rows, n_real * stride * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    if (n_int) CU_TRY(cudaMemcpyAsync(e->d_int[0].ptr, int_rows, n_int * stride * sizeof(int), cudaMemcpyHostToDevice, e->compute));
    const int grid = static_cast<int>(std::min<uint64_t>((n + kBlock - 1) / kBlock, static_cast<uint64_t>(e->sm_count) * 8));
    CU_TRY(vt->launch_replay(e->compute, grid, e->d_obs.ptr, static_cast<int>(n_obs), e->d_real[0].ptr, e->d_int[0].ptr, stride, n,
                             e->d_logw[0].ptr));
    ++e->launches;
    CU_TRY(cudaMemcpyAsync(logw_out, e->d_logw[0].ptr, n * sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    return 0;
}

int cpprob_sis_reduce_records(cpprob_sis_engine * e, const double * real_rows, int n_real, const int32_t * int_rows, int n_int,
                              const double * log_w, uint64_t stride, uint64_t n, cpprob_sis_stats * out)
{
    if (!e || !out || !log_w || n == 0 || stride < n || n_real < 0 || n_int < 0) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if ((n_real && !real_rows) || (n_int && !int_rows)) return fail(CPPROB_SIS_EINVAL, "null rows");
    if (int rc = use_device(e)) return rc;
    // int window from the data itself (host pass over the ints is cheap next to the upload)
    hist_window hw;
    if (n_int > 0) {
        int lo = std::numeric_limits<int>::max(), hi = std::numeric_limits<int>::min();
        for (int r = 0; r < n_int; ++r) {
            for (uint64_t i = 0; i < n; ++i) {
                const int v = int_rows[static_cast<uint64_t>(r) * stride + i];
                lo = std::min(lo, v);
                hi = std::max(hi, v);
            }
        }
        hw.lo = lo;
        const long long span = static_cast<long long>(hi) - lo + 1;
        if (span > 4096) return fail(CPPROB_SIS_ERANGE, "int predicts span more than 4096 values");
        hw.bins = static_cast<int>(span);
    }
    const int n_cols = kBaseCols + 2 * n_real + n_int * hw.bins;
    const unsigned n_chunks = static_cast<unsigned>((n + kSubChunk - 1) / kSubChunk);
    CU_TRY(e->d_real[0].reserve(std::max<size_t>(1, static_cast<size_t>(n_real) * stride)));
    CU_TRY(e->d_int[0].reserve(std::max<size_t>(1, static_cast<size_t>(n_int) * stride)));
    CU_TRY(e->d_logw[0].reserve(n));
    CU_TRY(e->d_w[0].reserve(n));
    CU_TRY(e->d_pilot.reserve(4));
    CU_TRY(e->d_partials.reserve(static_cast<size_t>(n_chunks) * n_cols));
    if (n_real) CU_TRY(cudaMemcpyAsync(e->d_real[0].ptr, real_rows, static_cast<size_t>(n_real) * stride * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    if (n_int) CU_TRY(cudaMemcpyAsync(e->d_int[0].ptr, int_rows, static_cast<size_t>(n_int) * stride * sizeof(int), cudaMemcpyHostToDevice, e->compute));
    CU_TRY(cudaMemcpyAsync(e->d_logw[0].ptr, log_w, n * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    uint64_t launches = 0;
    CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
    k_max_array<<<1, kBlock, 0, e->compute>>>(e->d_logw[0].ptr, n, e->d_pilot.ptr);
    CU_TRY(cudaGetLastError());
    k_row_base<<<n_chunks, kBlock, 0, e->compute>>>(e->d_logw[0].ptr, n, kSubChunk, e->d_pilot.ptr, e->d_w[0].ptr, nullptr, 0, 0, e->d_partials.ptr, n_cols);
    CU_TRY(cudaGetLastError());
    launches += 2;
    if (n_real > 0) {
        CU_TRY(launch_rows_moments(e->compute, n_chunks, e->d_real[0].ptr, e->d_w[0].ptr, stride, n, n_real, e->d_partials.ptr, n_cols));
        ++launches;
    }
    if (n_int > 0) {
        CU_TRY(launch_hist_all(e->compute, n_chunks, n_int, e->d_int[0].ptr, e->d_w[0].ptr, stride, n, hw, kBaseCols + 2 * n_real,
                               e->d_partials.ptr, n_cols, &launches));
    }
    CU_TRY(cudaEventRecord(e->ev_end, e->compute));
    double m_ref = 0.0;
    CU_TRY(cudaMemcpyAsync(&m_ref, e->d_pilot.ptr, sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    float ms1 = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms1, e->ev_begin, e->ev_end));
    double ms2 = 0.0;
    const int rc = merge_impl(e, e->d_partials.ptr, n_chunks, n_cols, n_real, n_int, hw, m_ref, n, out, &launches, &ms2);
    if (rc < 0) return rc;
    out->device_ms = ms1 + ms2;
    out->kernel_launches = launches;
    out->passes = 1;
    e->launches += launches;
    return 0;
}

// ---- device distribution layer -------------------------------------------------------------------
namespace {
int map_grid(const cpprob_sis_engine * e, uint64_t n)
{
    return static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>((n + kBlock - 1) / kBlock, static_cast<uint64_t>(e->sm_count) * 8)));
}
}  // namespace

int cpprob_sis_logpdf(cpprob_sis_engine * e, int kind, const double * params, int n_params, const double * x, uint64_t n, double * out)
{
    if (!e || !params || !x || !out || n_params < 0 || n_params > 8) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (n == 0) return 0;
    if (int rc = use_device(e)) return rc;
    CU_TRY(e->d_logw[0].reserve(n));
    CU_TRY(e->d_w[0].reserve(n));
    CU_TRY(cudaMemcpyAsync(e->d_logw[0].ptr, x, n * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    logpdf_op op;
    op.kind = kind;
    std::memset(&op.q, 0, sizeof op.q);
    for (int i = 0; i < n_params; ++i) op.q.p[i] = params[i];
    op.q.n = n_params;
    op.x = e->d_logw[0].ptr;
    op.out = e->d_w[0].ptr;
    k_map<<<map_grid(e, n), kBlock, 0, e->compute>>>(n, op);
    CU_TRY(cudaGetLastError());
    ++e->launches;
    CU_TRY(cudaMemcpyAsync(out, e->d_w[0].ptr, n * sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    return 0;
}

int cpprob_sis_sample(cpprob_sis_engine * e, int kind, const double * params, int n_params, uint64_t seed, uint64_t first_particle,
                      uint64_t n, double * out)
{
    if (!e || !params || !out || n_params < 0 || n_params > 8) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (n == 0) return 0;
    if (int rc = use_device(e)) return rc;
    CU_TRY(e->d_w[0].reserve(n));
    sample_op op;
    op.kind = kind;
    std::memset(&op.q, 0, sizeof op.q);
    for (int i = 0; i < n_params; ++i) op.q.p[i] = params[i];
    op.q.n = n_params;
    op.seed = seed;
    op.first = first_particle;
    op.out = e->d_w[0].ptr;
    CU_TRY(cudaFuncSetAttribute(k_map<sample_op>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(zig::kSharedBytes)));
    k_map<<<map_grid(e, n), kBlock, zig::kSharedBytes, e->compute>>>(n, op);
    CU_TRY(cudaGetLastError());
    ++e->launches;
    CU_TRY(cudaMemcpyAsync(out, e->d_w[0].ptr, n * sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    return 0;
}

int cpprob_sis_philox(cpprob_sis_engine * e, const uint32_t * ctr, const uint32_t * key, uint64_t n, uint32_t * out)
{
    if (!e || !ctr || !key || !out) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (n == 0) return 0;
    if (int rc = use_device(e)) return rc;
    // 4 + 2 + 4 words per block, staged in the double buffers (8-byte aligned)
    CU_TRY(e->d_logw[0].reserve(n * 2));
    CU_TRY(e->d_w[0].reserve(n * 2));
    CU_TRY(e->d_real[0].reserve(n));
    unsigned * d_ctr = reinterpret_cast<unsigned *>(e->d_logw[0].ptr);
    unsigned * d_out = reinterpret_cast<unsigned *>(e->d_w[0].ptr);
    unsigned * d_key = reinterpret_cast<unsigned *>(e->d_real[0].ptr);
    CU_TRY(cudaMemcpyAsync(d_ctr, ctr, n * 4 * sizeof(unsigned), cudaMemcpyHostToDevice, e->compute));
    CU_TRY(cudaMemcpyAsync(d_key, key, n * 2 * sizeof(unsigned), cudaMemcpyHostToDevice, e->compute));
    philox_op op{d_ctr, d_key, d_out};
    k_map<<<map_grid(e, n), kBlock, 0, e->compute>>>(n, op);
    CU_TRY(cudaGetLastError());
    ++e->launches;
    CU_TRY(cudaMemcpyAsync(out, d_out, n * 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    return 0;
}

int cpprob_sis_dmath(cpprob_sis_engine * e, int fn, const double * x, uint64_t n, double * out)
{
    if (!e || !x || !out) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (n == 0) return 0;
    if (int rc = use_device(e)) return rc;
    const uint64_t n_in = fn == 5 ? 2 * n : n;
    CU_TRY(e->d_logw[0].reserve(n_in));
    CU_TRY(e->d_w[0].reserve(n));
    CU_TRY(cudaMemcpyAsync(e->d_logw[0].ptr, x, n_in * sizeof(double), cudaMemcpyHostToDevice, e->compute));
    dmath_op op{fn, e->d_logw[0].ptr, e->d_w[0].ptr};
    k_map<<<map_grid(e, n), kBlock, 0, e->compute>>>(n, op);
    CU_TRY(cudaGetLastError());
    ++e->launches;
    CU_TRY(cudaMemcpyAsync(out, e->d_w[0].ptr, n * sizeof(double), cudaMemcpyDeviceToHost, e->compute));
    CU_TRY(cudaStreamSynchronize(e->compute));
    return 0;
}

// ---- roofline denominators -------------------------------------------------------------------------
int cpprob_sis_measure_dfma_peak(cpprob_sis_engine * e, double * tflops, double * sm_clock_mhz_est)
{
    if (!e || !tflops) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (int rc = use_device(e)) return rc;
    const int grid = e->sm_count * 8;
    CU_TRY(e->d_w[0].reserve(static_cast<size_t>(grid) * kBlock));
    const int iters = 4096;
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
        k_dfma_peak<<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        if (rep > 0) best_ms = std::min<double>(best_ms, ms);
        ++e->launches;
    }
    const double fmas = static_cast<double>(grid) * kBlock * iters * 32.0;
    *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    if (sm_clock_mhz_est) {
        // 64 DFMA lanes per SM per clock
        *sm_clock_mhz_est = fmas / (best_ms * 1e-3) / (64.0 * e->sm_count) / 1e6;
    }
    return 0;
}

// chains: independent DFMA chains per thread (1,2,4,8); blocks_per_sm CTAs of 256 threads per SM.
// Returns DFMA warp-instructions per cycle per SM sub-partition assuming `sm_mhz`.
int cpprob_sis_probe_dfma_chains(cpprob_sis_engine * e, int chains, int blocks_per_sm, double * ms_out, double * dfma_per_thread)
{
    if (!e || !ms_out) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (int rc = use_device(e)) return rc;
    const int grid = e->sm_count * blocks_per_sm;
    CU_TRY(e->d_w[0].reserve(static_cast<size_t>(grid) * kBlock));
    const int iters = 1024;
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
        switch (chains) {
        case 1: k_dfma_chains<1><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6); break;
        case 2: k_dfma_chains<2><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6); break;
        case 4: k_dfma_chains<4><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6); break;
        default: k_dfma_chains<8><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6); break;
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        if (rep > 0) best_ms = std::min<double>(best_ms, ms);
        ++e->launches;
    }
    *ms_out = best_ms;
    if (dfma_per_thread) *dfma_per_thread = static_cast<double>(iters) * 32.0 * (chains == 1 || chains == 2 || chains == 4 ? chains : 8);
    return 0;
}

int cpprob_sis_probe_issue(cpprob_sis_engine * e, int int_per_dfma, double * ms_out)
{
    if (!e || !ms_out || int_per_dfma < 0 || int_per_dfma > 3) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (int rc = use_device(e)) return rc;
    const int grid = e->sm_count * 8;
    CU_TRY(e->d_w[0].reserve(static_cast<size_t>(grid) * kBlock));
    const int iters = 2048;
    double best_ms = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
        switch (int_per_dfma) {
        case 0: k_issue_probe<0><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6, 0x9E3779B9u); break;
        case 1: k_issue_probe<1><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6, 0x9E3779B9u); break;
        case 2: k_issue_probe<2><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6, 0x9E3779B9u); break;
        default: k_issue_probe<3><<<grid, kBlock, 0, e->compute>>>(e->d_w[0].ptr, iters, 0.999999, 1.0e-6, 0x9E3779B9u); break;
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        if (rep > 0) best_ms = std::min<double>(best_ms, ms);
        ++e->launches;
    }
    *ms_out = best_ms;
    return 0;
}

int cpprob_sis_measure_store_peak(cpprob_sis_engine * e, double * gbytes_per_s)
{
    if (!e || !gbytes_per_s) return fail(CPPROB_SIS_EINVAL, "bad argument");
    if (int rc = use_device(e)) return rc;
    const size_t n = (1ull << 30) / sizeof(double);   // 1 GiB, well beyond L2
    CU_TRY(e->d_real[0].reserve(n));
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        CU_TRY(cudaEventRecord(e->ev_begin, e->compute));
        k_store_peak<<<e->sm_count * 16, kBlock, 0, e->compute>>>(reinterpret_cast<double2 *>(e->d_real[0].ptr), n / 2, 1.0 + rep);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(e->ev_end, e->compute));
        CU_TRY(cudaStreamSynchronize(e->compute));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
        if (rep > 0) best_ms = std::min<double>(best_ms, ms);
        ++e->launches;
    }
    *gbytes_per_s = static_cast<double>(n * sizeof(double)) / (best_ms * 1e-3) / 1e9;
    return 0;
}

// ---- posterior files ---------------------------------------------------------------------------------
namespace {
struct file_sink {
    cpprob_sis_engine * e;
    cpprob::text::posterior_writer * writer;
};
int file_sink_block(void * user, const cpprob_sis_block * blk)
{
    file_sink * fs = static_cast<file_sink *>(user);
    return fs->writer->append(*blk) ? 0 : 1;
}
}  // namespace

int cpprob_sis_infer_to_files(cpprob_sis_engine * e, int model_id, const double * obs, size_t n_obs, uint64_t n_particles,
                              const char * prefix, int emit, cpprob_sis_stats * out)
{
    if (!e || !out || !obs || !prefix) return fail(CPPROB_SIS_EINVAL, "null argument");
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    if (int rc = probe_structure(e, vt, obs, n_obs)) return rc;
    if (emit == CPPROB_SIS_EMIT_NONE) {
        shard_options none;
        if (int rc = run_full(e, vt, obs, n_obs, n_particles, none, out)) return rc;
        remove_record_files(prefix);
        if (!cpprob::text::write_ids(prefix, e->structure.ids)) return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".ids");
        if (!cpprob::text::write_stats_sidecar(prefix, *out, e->slots, e->structure.ids)) {
            return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".stats");
        }
        return 0;
    }
    cpprob::text::posterior_writer writer(prefix, e->slots);
    if (!writer.open()) return fail(CPPROB_SIS_EIO, std::string("cannot open posterior files for ") + prefix + ": " + std::strerror(errno));
    file_sink fs{e, &writer};
    shard_options so;
    so.emit = CPPROB_SIS_EMIT_ALL;
    // default: lines are formatted on the GPU and only text crosses PCIe; CPPROB_SIS_TEXT=host keeps the binary
    // rows -> pinned host -> std::to_chars path (same bytes, used as the cross-check in tests/test_files_gpu.py)
    const char * text_env = std::getenv("CPPROB_SIS_TEXT");
    if (text_env && std::strcmp(text_env, "host") == 0) {
        so.on_block = &file_sink_block;
        so.user = &fs;
    } else {
        so.text_writer = &writer;
    }
    const int rc = run_full(e, vt, obs, n_obs, n_particles, so, out);
    if (rc != 0) return rc;
    if (!writer.finish(e->structure.ids)) return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".ids");
    if (!cpprob::text::write_stats_sidecar(prefix, *out, e->slots, e->structure.ids)) {
        return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".stats");
    }
    return 0;
}

// Multi-GPU emission.  Rank r (= engines[r]) owns a contiguous range of global particle indices, so the posterior file of
// the run is the ranks' texts one after the other.  Pass A measures every rank's text (particle kernel + the length
// kernels, nothing copied or written); the files are extended ONCE to their final size; pass B runs the ranks
// concurrently, each formatting its own records on its own GPU and copying them into its own byte range of the shared
// mapping.  The estimators come from the ranks' partial rows, gathered on engines[0]'s GPU and merged as always.  Every
// byte of every file equals what one GPU writes for the same seed.
int cpprob_sis_infer_to_files_multi(cpprob_sis_engine * const * engines, int n_engines, int model_id, const double * obs, size_t n_obs,
                                    uint64_t n_particles, const char * prefix, cpprob_sis_stats * out)
{
    if (!engines || n_engines <= 0 || !obs || !prefix || !out) return fail(CPPROB_SIS_EINVAL, "bad argument");
    for (int r = 0; r < n_engines; ++r) {
        if (!engines[r]) return fail(CPPROB_SIS_EINVAL, "null engine");
        if (engines[r]->seed != engines[0]->seed) return fail(CPPROB_SIS_EINVAL, "all engines of a multi-GPU run must share one seed");
    }
    const cpprob_sis_model_vtable * vt = model_of(model_id);
    if (!vt) return fail(CPPROB_SIS_ENOMODEL, "unknown model id");
    cpprob_sis_engine * primary = engines[0];
    if (int rc = probe_structure(primary, vt, obs, n_obs)) return rc;
    const size_t R = static_cast<size_t>(n_engines);
    const int n_real = static_cast<int>(primary->structure.n_real), n_int = static_cast<int>(primary->structure.n_int);

    std::vector<std::unique_ptr<cpprob::text::posterior_writer>> writers;
    for (size_t r = 0; r < R; ++r) writers.emplace_back(new cpprob::text::posterior_writer(prefix, primary->slots));
    std::vector<shard_result> res(R);
    // one host thread per rank: a shard that emits waits for its own copies and writes its own file range
    auto run_all = [&](const shard_options & proto, const double * mo, const hist_window * ho) -> int {
        std::vector<int> rc(R, 0);
        std::vector<std::string> msg(R);
        std::vector<std::thread> pool;
        for (size_t r = 0; r < R; ++r) {
            pool.emplace_back([&, r] {
                shard_options so = proto;
                if (so.text_writer) so.text_writer = writers[r].get();
                res[r] = shard_result();
                rc[r] = run_shard_impl(engines[r], vt, obs, n_obs, n_particles, static_cast<int>(r), n_engines, mo, ho, so, &res[r]);
                if (rc[r] != 0) msg[r] = g_last_error;                 // thread-local: carry it out
            });
        }
        for (auto & t : pool) t.join();
        for (size_t r = 0; r < R; ++r) {
            if (rc[r] != 0) return fail(rc[r], "device " + std::to_string(engines[r]->device) + ": " + msg[r]);
        }
        return 0;
    };
    // the ranks' partial rows -> engines[0]'s GPU in rank order -> merge
    uint64_t launches = 0;
    double total_ms = 0.0;
    auto gather_merge = [&]() -> int {
        if (int rc = use_device(primary)) return rc;
        const int n_cols = res[0].n_cols;
        const uint32_t rows_total = res[0].n_rows_total;
        CU_TRY(primary->d_gather.reserve(static_cast<size_t>(rows_total) * n_cols));
        double shard_ms = 0.0;
        for (size_t r = 0; r < R; ++r) {
            shard_ms = std::max(shard_ms, res[r].device_ms);
            launches += res[r].launches;
            if (res[r].n_rows_local == 0) continue;
            CU_TRY(cudaMemcpyPeerAsync(primary->d_gather.ptr + static_cast<size_t>(res[r].row_first) * n_cols, primary->device, res[r].rows,
                                       engines[r]->device, static_cast<size_t>(res[r].n_rows_local) * n_cols * sizeof(double), primary->compute));
        }
        double merge_ms = 0.0;
        const int mrc = merge_impl(primary, primary->d_gather.ptr, rows_total, n_cols, n_real, n_int, res[0].hw, res[0].m_ref, n_particles, out,
                                   &launches, &merge_ms);
        total_ms += shard_ms + merge_ms;
        return mrc;
    };

    // pass A: sizes
    shard_options so;
    so.emit = CPPROB_SIS_EMIT_ALL;
    so.text_writer = writers[0].get();
    so.count_text_only = true;
    if (int rc = run_all(so, nullptr, nullptr)) return rc;
    std::vector<std::array<unsigned long long, 2>> sizes(R);
    unsigned long long total[2] = {0, 0};
    for (size_t r = 0; r < R; ++r) {
        for (int k = 0; k < 2; ++k) { sizes[r][static_cast<size_t>(k)] = res[r].text_total[k]; total[k] += res[r].text_total[k]; }
    }
    // the files grow once; every rank learns where its records start
    off_t begin[2] = {0, 0};
    if (n_real > 0) {
        begin[0] = cpprob::text::posterior_writer::extend_file(std::string(prefix) + ".real", total[0]);
        if (begin[0] < 0) return fail(CPPROB_SIS_EIO, std::string("cannot extend ") + prefix + ".real: " + std::strerror(errno));
    }
    if (n_int > 0) {
        begin[1] = cpprob::text::posterior_writer::extend_file(std::string(prefix) + ".int", total[1]);
        if (begin[1] < 0) return fail(CPPROB_SIS_EIO, std::string("cannot extend ") + prefix + ".int: " + std::strerror(errno));
    }
    off_t at[2] = {begin[0], begin[1]};
    for (size_t r = 0; r < R; ++r) {
        if (!writers[r]->open_at(at[0], at[1])) return fail(CPPROB_SIS_EIO, std::string("cannot open posterior files for ") + prefix + ": " + std::strerror(errno));
        at[0] += static_cast<off_t>(sizes[r][0]);
        at[1] += static_cast<off_t>(sizes[r][1]);
    }
    // pass B: records and partial sums
    so.count_text_only = false;
    if (int rc = run_all(so, nullptr, nullptr)) return rc;
    int mrc = gather_merge();
    if (mrc < 0) return mrc;
    // re-basing / a wider histogram window only redo the sums (the records are written), as on one GPU
    double m_ref_override = 0.0;
    hist_window hw_override;
    int passes = 1;
    for (; passes < 3; ++passes) {
        const double * mo = nullptr;
        const hist_window * ho = nullptr;
        if (mrc == 1) { m_ref_override = out->max_log_w; mo = &m_ref_override; }
        hist_window hw = res[0].hw;
        if (vt->int_states > 0 && n_int > 0 && out->sums[col::int_oor] != 0.0) {
            return fail(CPPROB_SIS_ERANGE, std::string("model ") + vt->name + " predicted an integral value outside [0, " + std::to_string(vt->int_states) +
                                               "), the range it declares (int_predict_states)");
        }
        if (widen_window(*out, n_int, &hw)) {
            if (hw.bins > 4096) return fail(CPPROB_SIS_ERANGE, "int predicts span more than 4096 values");
            hw_override = hw;
            ho = &hw_override;
        }
        if (!mo && !ho) break;
        shard_options again;
        again.force_rows = 1;
        if (int rc = run_all(again, mo, ho)) return rc;
        mrc = gather_merge();
        if (mrc < 0) return mrc;
    }
    out->device_ms = total_ms;
    out->kernel_launches = launches;
    out->passes = passes;
    out->path = CPPROB_SIS_PATH_ROWS;
    primary->launches += launches;
    for (size_t r = 1; r < R; ++r) writers[r].reset();                 // closes their descriptors
    if (!writers[0]->finish(primary->structure.ids)) return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".ids");
    if (!cpprob::text::write_stats_sidecar(prefix, *out, primary->slots, primary->structure.ids)) {
        return fail(CPPROB_SIS_EIO, std::string("cannot write ") + prefix + ".stats");
    }
    return 0;
}

int cpprob_sis_text_stage_stats(cpprob_sis_engine * e, double * kernel_ms, double * copy_ms, double * write_s, uint64_t * bytes,
                                uint64_t * fixups)
{
    if (!e) return fail(CPPROB_SIS_EINVAL, "null engine");
    if (kernel_ms) *kernel_ms = e->text_kernel_ms;
    if (copy_ms) *copy_ms = e->text_copy_ms;
    if (write_s) *write_s = e->text_write_s;
    if (bytes) *bytes = e->text_bytes;
    if (fixups) *fixups = e->text_fixups;
    return 0;
}

}  // extern "C"
