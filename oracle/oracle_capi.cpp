// ORACLE — TEST INFRASTRUCTURE ONLY (see cpprob_oracle.hpp).  C entry points for ctypes.
#include <chrono>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "cpprob_oracle.hpp"

using namespace oracle;

namespace {
std::string g_text;

template<class F>
double timed_inference(const F & model, std::size_t n, const char * prefix, int how)
{
    const auto t0 = std::chrono::steady_clock::now();
    inference(StateType::sis, model, n, prefix, how == 0 ? flavour::faithful : flavour::fast);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
}  // namespace

extern "C" {

// kind ids are those of include/cpprob_sis.h (CPPROB_SIS_DIST_*): 0 normal 1 uniform_real
// 2 uniform_smallint 3 discrete 4 poisson
int oracle_logpdf(int kind, const double * q, int nq, const double * x, unsigned long long n, double * out)
{
    for (unsigned long long i = 0; i < n; ++i) {
        switch (kind) {
        case 0: out[i] = logpdf<normal_distribution<>>()(normal_distribution<>(q[0], q[1]), x[i]); break;
        case 1: out[i] = logpdf<uniform_real_distribution<>>()(uniform_real_distribution<>(q[0], q[1]), x[i]); break;
        case 2:
            out[i] = logpdf<uniform_smallint<long long>>()(uniform_smallint<long long>(static_cast<long long>(q[0]), static_cast<long long>(q[1])),
                                                           static_cast<long long>(x[i]));
            break;
        case 3:
            out[i] = logpdf<discrete_distribution<long long>>()(discrete_distribution<long long>(q, q + nq), static_cast<long long>(x[i]));
            break;
        case 4: out[i] = logpdf<poisson_distribution<long long>>()(poisson_distribution<long long>(q[0]), static_cast<long long>(x[i])); break;
        default: return -1;
        }
    }
    return 0;
}

// Runs the restated cpprob::inference(StateType::sis, ...) and returns the wall time in seconds
// (< 0 on error).  how: 0 = faithful (3 file appends per trace), 1 = fast (buffered streams).
// progress: 0 none, 1 to stdout (as the reference), 2 to /dev/null (same formatting + flush work).
double oracle_run(const char * model, const double * obs, int n_obs, unsigned long long n, const char * prefix, int how,
                  unsigned seed, int progress)
{
    seed_rng(seed);
    engine & e = engine::get();
    e.replay_values = nullptr;
    static std::ofstream devnull;
    if (progress == 2 && !devnull.is_open()) devnull.open("/dev/null");
    e.progress = progress == 1 ? &std::cout : (progress == 2 ? static_cast<std::ostream *>(&devnull) : nullptr);
    const std::string m(model);
    const std::vector<double> o(obs, obs + n_obs);
    if (m == "gaussian_unknown_mean" && n_obs == 2) return timed_inference([&] { models::gaussian_unknown_mean(o[0], o[1]); }, n, prefix, how);
    if (m == "gaussian_unknown_mean_mu" && n_obs == 2) return timed_inference([&] { models::gaussian_unknown_mean_mu(o[0], o[1]); }, n, prefix, how);
    if (m == "linear_gaussian_1d") return timed_inference([&] { models::linear_gaussian_1d(o); }, n, prefix, how);
    if (m == "hmm") return timed_inference([&] { models::hmm(o); }, n, prefix, how);
    if (m == "gaussian_2d_unk_mean") return timed_inference([&] { models::gaussian_2d_unk_mean(o); }, n, prefix, how);
    if (m == "normal_rejection_sampling" && n_obs == 2) return timed_inference([&] { models::normal_rejection_sampling(o[0], o[1]); }, n, prefix, how);
    if (m.rfind("poly_adjustment_", 0) == 0) return timed_inference([&] { models::poly_adjustment(m.back() - '0', o); }, n, prefix, how);
    if (m == "linear_regression") return timed_inference([&] { models::linear_regression(o); }, n, prefix, how);
    if (m == "all_distr") return timed_inference([&] { models::all_distr(0, 0); }, n, prefix, how);
    return -1.0;
}

// The restated inference loop on PRESCRIBED sampled values (`values`: program order, trace after trace): writes
// <prefix>.real / .int / .ids exactly as a sampling run would.  Counterpart of oracle/ref_sis.cpp --replay, which runs the
// reference's own loop on the same values (tests/test_ref_sis.py compares the files byte for byte).
int oracle_replay_files(const char * model, const double * obs, int n_obs, const double * values, unsigned long long n_values,
                        unsigned long long n_traces, const char * prefix, int how)
{
    engine & e = engine::get();
    e.progress = nullptr;
    e.replay_values = values;
    e.replay_pos = 0;
    const std::string m(model);
    const std::vector<double> o(obs, obs + n_obs);
    double s = -1.0;
    if (m == "gaussian_unknown_mean" && n_obs == 2) s = timed_inference([&] { models::gaussian_unknown_mean(o[0], o[1]); }, n_traces, prefix, how);
    else if (m == "gaussian_unknown_mean_mu" && n_obs == 2) s = timed_inference([&] { models::gaussian_unknown_mean_mu(o[0], o[1]); }, n_traces, prefix, how);
    else if (m == "linear_gaussian_1d") s = timed_inference([&] { models::linear_gaussian_1d(o); }, n_traces, prefix, how);
    else if (m == "hmm") s = timed_inference([&] { models::hmm(o); }, n_traces, prefix, how);
    else if (m == "gaussian_2d_unk_mean") s = timed_inference([&] { models::gaussian_2d_unk_mean(o); }, n_traces, prefix, how);
    else if (m.rfind("poly_adjustment_", 0) == 0) s = timed_inference([&] { models::poly_adjustment(m.back() - '0', o); }, n_traces, prefix, how);
    else if (m == "normal_rejection_sampling" && n_obs == 2) s = timed_inference([&] { models::normal_rejection_sampling(o[0], o[1]); }, n_traces, prefix, how);
    else if (m == "all_distr") s = timed_inference([&] { models::all_distr(0, 0); }, n_traces, prefix, how);
    else if (m == "linear_regression") s = timed_inference([&] { models::linear_regression(o); }, n_traces, prefix, how);
    const bool all_used = e.replay_pos == n_values;
    e.replay_values = nullptr;
    return s < 0 ? -1 : (all_used ? 0 : -2);
}

// log_w of one trace whose sample statements return `values` in order (replay gate).
int oracle_replay_logw(const char * model, const double * obs, int n_obs, const double * values, unsigned long long n_traces,
                       int values_per_trace, double * logw_out)
{
    engine & e = engine::get();
    const std::string m(model);
    const std::vector<double> o(obs, obs + n_obs);
    for (unsigned long long t = 0; t < n_traces; ++t) {
        e.start_trace();
        e.replay_values = values + t * static_cast<unsigned long long>(values_per_trace);
        e.replay_pos = 0;
        if (m == "gaussian_unknown_mean") models::gaussian_unknown_mean(o[0], o[1]);
        else if (m == "gaussian_unknown_mean_mu") models::gaussian_unknown_mean_mu(o[0], o[1]);
        else if (m == "linear_gaussian_1d") models::linear_gaussian_1d(o);
        else if (m == "hmm") models::hmm(o);
        else if (m == "gaussian_2d_unk_mean") models::gaussian_2d_unk_mean(o);
        else if (m.rfind("poly_adjustment_", 0) == 0) models::poly_adjustment(m.back() - '0', o);
        else if (m == "linear_regression") models::linear_regression(o);
        else { e.replay_values = nullptr; return -1; }
        logw_out[t] = e.trace.log_w;
    }
    e.replay_values = nullptr;
    e.ids.clear();
    return 0;
}

// StatsPrinter{prefix} streamed to a string (exact console text of the reference's post-processing).
const char * oracle_stats_text(const char * prefix)
{
    std::ostringstream os;
    os << stats_printer{prefix} << std::endl;
    g_text = os.str();
    return g_text.c_str();
}

// Numeric estimators.  Rows are ordered by (id, k, component); a vector-valued predict contributes one row per
// component.  Returns the number of rows written, or -1.  For ints use oracle_stats_int.
int oracle_stats_real(const char * prefix, int max_rows, int * ids, int * ks, double * mean, double * var)
{
    stats_printer sp{prefix};
    int r = 0;
    for (const auto & kv : sp.reals()) {
        int k = 0;
        for (const auto & d : kv.second) {
            const nd_value m = d.mean();
            const nd_value v = d.variance(m);
            for (std::size_t c = 0; c < m.v.size(); ++c) {
                if (r >= max_rows) return -1;
                ids[r] = static_cast<int>(kv.first);
                ks[r] = k;
                mean[r] = m.v[c];
                var[r] = v.v[c];
                ++r;
            }
            ++k;
        }
    }
    return r;
}

// probabilities of values lo..lo+bins-1 per (id,k) row, MAP and num points
int oracle_stats_int(const char * prefix, int max_rows, int lo, int bins, int * ids, int * ks, double * prob, int * map_out,
                     unsigned long long * num_points)
{
    stats_printer sp{prefix};
    int r = 0;
    for (const auto & kv : sp.ints()) {
        int k = 0;
        for (const auto & d : kv.second) {
            if (r >= max_rows) return -1;
            ids[r] = static_cast<int>(kv.first);
            ks[r] = k++;
            const auto distr = d.distribution();
            for (int b = 0; b < bins; ++b) {
                const auto it = distr.find(lo + b);
                prob[r * bins + b] = it == distr.end() ? 0.0 : it->second;
            }
            map_out[r] = d.max_a_posteriori(distr);
            num_points[r] = d.num_points();
            ++r;
        }
    }
    return r;
}

// Parses a posterior record file with the reference grammar into dense arrays.
// kind 0: real file, kind 1: int file.  values is [n_records][per_record]; returns n_records or -1.
long long oracle_parse_records(const char * path, int kind, int per_record, unsigned long long max_records, int * ids_out,
                               double * values, double * logw)
{
    std::ifstream f(path);
    if (!f.is_open()) return -1;
    unsigned long long n = 0;
    for (std::string line; std::getline(f, line);) {
        if (n >= max_records) return -1;
        std::istringstream iss(line);
        if (kind == 0) {
            // values are flattened component by component: per_record counts doubles, not predicts
            std::pair<std::vector<std::pair<std::size_t, nd_value>>, double> rec;
            if (!(iss >> rec)) return -1;
            int j = 0;
            for (const auto & el : rec.first) {
                for (double x : el.second.v) {
                    if (j >= per_record) return -1;
                    values[n * per_record + j] = x;
                    if (n == 0) ids_out[j] = static_cast<int>(el.first);
                    ++j;
                }
            }
            if (j != per_record) return -1;
            logw[n] = rec.second;
        } else {
            std::pair<std::vector<std::pair<std::size_t, int>>, double> rec;
            if (!(iss >> rec) || static_cast<int>(rec.first.size()) != per_record) return -1;
            for (int j = 0; j < per_record; ++j) {
                values[n * per_record + j] = rec.first[j].second;
                if (n == 0) ids_out[j] = static_cast<int>(rec.first[j].first);
            }
            logw[n] = rec.second;
        }
        ++n;
    }
    return static_cast<long long>(n);
}

void oracle_philox(const unsigned * ctr, const unsigned * key, unsigned long long n, unsigned * out)
{
    for (unsigned long long i = 0; i < n; ++i) philox4x32_10(ctr + 4 * i, key + 2 * i, out + 4 * i);
}

}  // extern "C"
