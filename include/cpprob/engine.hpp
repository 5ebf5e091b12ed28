// cpprob-b200: C++14 RAII face of the C ABI (include/cpprob_sis.h).  Errors that the ABI reports as
// negative codes are re-thrown as std::runtime_error, the reference's error style
// (/root/reference include/cpprob/socket.hpp:70,75; include/cpprob/distributions/truncated.hpp:101).
#ifndef CPPROB_ENGINE_HPP
#define CPPROB_ENGINE_HPP

#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "cpprob_sis.h"

namespace cpprob {
namespace sis {

inline void check(int rc, const char * what)
{
    if (rc < 0) {
        throw std::runtime_error(std::string(what) + ": " + cpprob_sis_last_error());
    }
}

// Seed policy: CPPROB_SIS_SEED if set (reproducible runs), otherwise std::random_device — the
// reference's generator is seeded from random_device too (src/cpprob/utils.cpp:16-20).
inline std::uint64_t default_seed()
{
    if (const char * s = std::getenv("CPPROB_SIS_SEED")) {
        return std::strtoull(s, nullptr, 0);
    }
    std::random_device rd;
    return (static_cast<std::uint64_t>(rd()) << 32) | rd();
}

inline int default_device()
{
    if (const char * s = std::getenv("CPPROB_SIS_DEVICE")) return std::atoi(s);
    return 0;
}

// CPPROB_SIS_DEVICES="0,1,2,3": the GPUs a stats-only inference is sharded over (empty: one GPU)
inline std::vector<int> device_list()
{
    std::vector<int> out;
    if (const char * s = std::getenv("CPPROB_SIS_DEVICES")) {
        const char * p = s;
        while (*p) {
            char * end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            out.push_back(static_cast<int>(v));
            p = (*end == ',') ? end + 1 : end;
        }
    }
    return out;
}

class engine {
public:
    explicit engine(int device = default_device(), std::uint64_t seed = default_seed())
    {
        cpprob_sis_config cfg;
        cfg.device = device;
        cfg.seed = seed;
        cfg.max_batch = 0;
        cfg.blocks_per_sm = 0;
        check(cpprob_sis_create(&cfg, &h_), "cpprob_sis_create");
    }
    ~engine() { cpprob_sis_destroy(h_); }
    engine(const engine &) = delete;
    engine & operator=(const engine &) = delete;

    cpprob_sis_engine * handle() const { return h_; }

    void set_seed(std::uint64_t seed) { check(cpprob_sis_set_seed(h_, seed), "cpprob_sis_set_seed"); }

    int model_id(const std::string & name) const
    {
        const int id = cpprob_sis_find_model(name.c_str());
        check(id, "cpprob_sis_find_model");
        return id;
    }

    cpprob_sis_stats infer_to_files(int model, const std::vector<double> & obs, std::uint64_t n, const std::string & prefix, int emit)
    {
        cpprob_sis_stats st;
        check(cpprob_sis_infer_to_files(h_, model, obs.data(), obs.size(), n, prefix.c_str(), emit, &st), "cpprob_sis_infer_to_files");
        return st;
    }

    // stats-only inference sharded over several engines (this one is rank 0 and owns the results)
    cpprob_sis_stats run_multi(const std::vector<engine *> & others, int model, const std::vector<double> & obs, std::uint64_t n)
    {
        std::vector<cpprob_sis_engine *> hs(1, h_);
        for (engine * e : others) hs.push_back(e->handle());
        cpprob_sis_stats st;
        check(cpprob_sis_run_multi(hs.data(), static_cast<int>(hs.size()), model, obs.data(), obs.size(), n, &st), "cpprob_sis_run_multi");
        return st;
    }

    // posterior files written by several engines (this one is rank 0 and owns the results)
    cpprob_sis_stats infer_to_files_multi(const std::vector<engine *> & others, int model, const std::vector<double> & obs, std::uint64_t n,
                                          const std::string & prefix)
    {
        std::vector<cpprob_sis_engine *> hs(1, h_);
        for (engine * e : others) hs.push_back(e->handle());
        cpprob_sis_stats st;
        check(cpprob_sis_infer_to_files_multi(hs.data(), static_cast<int>(hs.size()), model, obs.data(), obs.size(), n, prefix.c_str(), &st),
              "cpprob_sis_infer_to_files_multi");
        return st;
    }

    cpprob_sis_stats run(int model, const std::vector<double> & obs, std::uint64_t n)
    {
        cpprob_sis_stats st;
        check(cpprob_sis_run(h_, model, obs.data(), obs.size(), n, nullptr, &st), "cpprob_sis_run");
        return st;
    }

    cpprob_sis_stats reduce_records(const double * real_rows, int n_real, const std::int32_t * int_rows, int n_int,
                                    const double * log_w, std::uint64_t stride, std::uint64_t n)
    {
        cpprob_sis_stats st;
        check(cpprob_sis_reduce_records(h_, real_rows, n_real, int_rows, n_int, log_w, stride, n, &st), "cpprob_sis_reduce_records");
        return st;
    }

private:
    cpprob_sis_engine * h_ = nullptr;
};

// One engine per GPU for the whole process, re-seeded per call: what cpprob::inference uses.  The reference's
// inference() has no per-call set-up to speak of (cpprob.hpp:184-192 resets a few statics); creating a CUDA context,
// streams, events and device tables on every call would put milliseconds in front of a 10,000-particle run whose
// kernels take microseconds.  Engines are released when the process ends.
inline engine & cached_engine(int device)
{
    static std::mutex m;
    static std::map<int, std::unique_ptr<engine>> cache;
    std::lock_guard<std::mutex> lock(m);
    std::unique_ptr<engine> & slot = cache[device];
    if (!slot) slot.reset(new engine(device, 0));
    return *slot;
}

}  // namespace sis
}  // namespace cpprob
#endif  // CPPROB_ENGINE_HPP
