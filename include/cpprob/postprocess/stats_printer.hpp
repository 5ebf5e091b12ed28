// cpprob-b200: cpprob::StatsPrinter — posterior estimators printed in the reference's format, with the
// estimator arithmetic done on the GPU.
//
// Reference: /root/reference include/cpprob/postprocess/stats_printer.hpp:22-121 and
// empirical_distribution.hpp:16-147.  `StatsPrinter{path}` reads <path>.ids and the record files
// <path>.int / <path>.real (grammar of serialization.hpp), groups the predicted values by
// (address id, k = occurrence index of that id within the record) (:98,106-118) and prints, per id,
//     Estimators for <path>.real            Estimators for <path>.int
//     <name>[ k]:                           <name>[ k]:
//       Mean: m                               Distribution:
//       Variance: v                             v: p
//                                             MAP: v
//                                             Num points: n
// (`k` only when the id occurs more than once per record, :49-51; the header line repeated per id,
// :44-45,59-60; numbers with the stream's default 6 significant digits).
//
// Here the records are parsed on the host into SoA rows and handed to cpprob_sis_reduce_records: the
// max / log-sum-exp / weighted moment / weighted histogram passes that EmpiricalDistribution runs
// three times per statistic on the CPU are one pass of the engine's reduction kernels.  When the
// record files are absent (inference ran with CPPROB_SIS_EMIT=none) the estimators come from the
// <path>.stats sidecar the engine wrote.  Extra accessors expose what the reference does not compute:
// ESS, log-evidence, max log-weight.
#ifndef CPPROB_STATS_PRINTER_HPP
#define CPPROB_STATS_PRINTER_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "cpprob/engine.hpp"

namespace cpprob {

class StatsPrinter {
public:
    struct real_estimate { std::vector<double> mean, variance; };   // one element per component (1 for a scalar predict)
    struct int_estimate {
        std::map<int, double> distribution;
        int map = 0;
        std::size_t num_points = 0;
    };

    StatsPrinter(const std::string & file_path) : file_name_{file_path}
    {
        const std::string file_ids = file_path + ".ids";
        std::ifstream ids_file(file_ids.c_str());
        if (!ids_file.is_open()) {
            std::cerr << file_ids << " not found." << std::endl;
            return;
        }
        for (std::string line; std::getline(ids_file, line);) {
            ids_.emplace_back(std::move(line));
        }
        const bool have_int = load_records(file_path + ".int", true);
        const bool have_real = load_records(file_path + ".real", false);
        if (!have_int && !have_real) load_sidecar(file_path + ".stats");
    }

    friend std::ostream & operator<<(std::ostream & out, const StatsPrinter & sp)
    {
        for (const auto & kv : sp.real_) {
            out << "Estimators for " << sp.file_name_ << ".real" << std::endl;
            std::size_t i = 0;
            for (const auto & est : kv.second) {
                out << sp.ids_[kv.first];
                if (kv.second.size() > 1) {
                    out << ' ' << i;
                }
                out << ':' << std::endl;
                out << "  Mean: ";
                print_value(out, est.mean);
                out << std::endl << "  Variance: ";
                print_value(out, est.variance);
                out << std::endl;
                ++i;
            }
        }
        for (const auto & kv : sp.int_) {
            out << "Estimators for " << sp.file_name_ << ".int" << std::endl;
            std::size_t i = 0;
            for (const auto & est : kv.second) {
                out << sp.ids_[kv.first];
                if (kv.second.size() > 1) {
                    out << ' ' << i;
                }
                out << ':' << std::endl
                    << "  Distribution:\n";
                for (const auto & x_w : est.distribution) {
                    out << "    " << x_w.first << ": " << x_w.second << std::endl;
                }
                out << "  MAP: " << est.map << std::endl;
                out << "  Num points: " << est.num_points << std::endl;
                ++i;
            }
        }
        return out;
    }

    // (id -> estimates by occurrence index k)
    const std::map<std::size_t, std::vector<real_estimate>> & real_estimates() const { return real_; }
    const std::map<std::size_t, std::vector<int_estimate>> & int_estimates() const { return int_; }
    const std::vector<std::string> & ids() const { return ids_; }
    // additions of the on-device reduction (NaN if unknown)
    double ess() const { return ess_; }
    double log_evidence() const { return log_evidence_; }
    double max_log_weight() const { return max_log_w_; }

private:
    // NDArray printing (ndarray.hpp:273-288): a scalar prints bare, a vector as [a b c]
    static void print_value(std::ostream & out, const std::vector<double> & v)
    {
        if (v.size() == 1) { out << v[0]; return; }
        out << '[';
        for (std::size_t i = 0; i < v.size(); ++i) out << (i ? " " : "") << v[i];
        out << ']';
    }

    struct col_key {
        std::size_t id, k;
        int comp, width;
        bool operator<(const col_key & o) const { return id != o.id ? id < o.id : (k != o.k ? k < o.k : comp < o.comp); }
    };

    // `([(id v) (id v) ...] logw)` — hand-rolled scanner for the grammar of serialization.hpp:41-98
    struct item { std::size_t id; int comp; int width; double v; };
    static bool parse_line(const char * p, bool is_int, std::vector<item> & items, double & logw)
    {
        auto skip = [&p] { while (*p == ' ' || *p == '\t' || *p == '\r') ++p; };
        items.clear();
        skip();
        if (*p++ != '(') return false;
        skip();
        if (*p++ != '[') return false;
        for (;;) {
            skip();
            if (*p == ']') { ++p; break; }
            if (*p++ != '(') return false;
            char * end = nullptr;
            const unsigned long long id = std::strtoull(p, &end, 10);
            if (end == p) return false;
            p = end;
            skip();
            if (!is_int && *p == '[') {                    // vector value: (id [v0 v1 ...])
                ++p;
                const std::size_t first = items.size();
                for (int comp = 0;; ++comp) {
                    skip();
                    if (*p == ']') { ++p; break; }
                    const double v = std::strtod(p, &end);
                    if (end == p) return false;
                    p = end;
                    items.push_back(item{static_cast<std::size_t>(id), comp, 0, v});
                }
                for (std::size_t j = first; j < items.size(); ++j) items[j].width = static_cast<int>(items.size() - first);
            } else {
                const double v = is_int ? static_cast<double>(std::strtol(p, &end, 10)) : std::strtod(p, &end);
                if (end == p) return false;
                p = end;
                items.push_back(item{static_cast<std::size_t>(id), 0, 1, v});
            }
            skip();
            if (*p++ != ')') return false;
        }
        char * end = nullptr;
        logw = std::strtod(p, &end);
        if (end == p) return false;
        p = end;
        skip();
        return *p == ')';
    }

    bool load_records(const std::string & path, bool is_int)
    {
        std::ifstream file(path.c_str());
        if (!file.is_open()) return false;

        // column-major staging: one growing vector per (id, k) column
        std::map<col_key, std::size_t> col_of;
        std::vector<col_key> keys;
        std::vector<std::vector<double>> cols;
        std::vector<std::vector<double>> col_logw;   // only used if the records are ragged
        std::vector<double> log_w;
        std::vector<item> items;
        bool ragged = false;
        std::size_t n = 0;
        for (std::string line; std::getline(file, line);) {
            double lw;
            if (!parse_line(line.c_str(), is_int, items, lw)) {
                std::cerr << "Bad format in line:\n" << line << std::endl;
                std::exit(EXIT_FAILURE);
            }
            std::map<std::size_t, std::size_t> seen;   // occurrences of each id within this record
            for (const auto & it : items) {
                if (it.comp == 0) ++seen[it.id];
                const col_key key{it.id, seen[it.id] - 1, it.comp, it.width};
                auto found = col_of.find(key);
                if (found == col_of.end()) {
                    found = col_of.emplace(key, cols.size()).first;
                    keys.push_back(key);
                    cols.emplace_back();
                    col_logw.emplace_back();
                    if (n != 0) ragged = true;
                }
                cols[found->second].push_back(it.v);
                col_logw[found->second].push_back(lw);
            }
            log_w.push_back(lw);
            ++n;
            if (!ragged) {
                for (const auto & c : cols) if (c.size() != n) { ragged = true; break; }
            }
        }
        if (n == 0 || cols.empty()) return true;

        sis::engine & engine = sis::cached_engine(sis::default_device());      // shared with cpprob::inference
        if (!ragged) {
            reduce_dense(engine, is_int, keys, cols, log_w, n);
        } else {
            // records of differing shape: every (id, k) has its own weight vector, as in the reference
            for (std::size_t c = 0; c < cols.size(); ++c) {
                std::vector<std::vector<double>> one(1, cols[c]);
                std::vector<col_key> key(1, keys[c]);
                reduce_dense(engine, is_int, key, one, col_logw[c], cols[c].size());
            }
        }
        return true;
    }

    void reduce_dense(sis::engine & engine, bool is_int, const std::vector<col_key> & keys,
                      const std::vector<std::vector<double>> & cols, const std::vector<double> & log_w, std::size_t n)
    {
        const int rows = static_cast<int>(cols.size());
        cpprob_sis_stats st;
        if (is_int) {
            std::vector<std::int32_t> flat(static_cast<std::size_t>(rows) * n);
            for (int r = 0; r < rows; ++r) for (std::size_t i = 0; i < n; ++i) flat[r * n + i] = static_cast<std::int32_t>(cols[r][i]);
            st = engine.reduce_records(nullptr, 0, flat.data(), rows, log_w.data(), n, n);
        } else {
            std::vector<double> flat(static_cast<std::size_t>(rows) * n);
            for (int r = 0; r < rows; ++r) std::memcpy(&flat[r * n], cols[r].data(), n * sizeof(double));
            st = engine.reduce_records(flat.data(), rows, nullptr, 0, log_w.data(), n, n);
        }
        ess_ = st.ess;
        log_evidence_ = st.log_evidence;
        max_log_w_ = st.max_log_w;
        for (int r = 0; r < rows; ++r) {
            const std::size_t id = keys[r].id, k = keys[r].k;
            if (is_int) {
                auto & vec = int_[id];
                if (vec.size() <= k) vec.resize(k + 1);
                int_estimate & e = vec[k];
                for (int b = 0; b < st.int_bins; ++b) {
                    const double p = st.int_prob[static_cast<std::size_t>(r) * st.int_bins + b];
                    // the reference's map only has the values that occur (empirical_distribution.hpp:36-38)
                    if (p > 0.0 || value_occurs(cols[r], static_cast<double>(st.int_lo + b))) e.distribution[static_cast<int>(st.int_lo + b)] = p;
                }
                e.map = static_cast<int>(st.int_map[r]);
                e.num_points = n;
            } else {
                auto & vec = real_[id];
                if (vec.size() <= k) vec.resize(k + 1);
                if (vec[k].mean.size() < static_cast<std::size_t>(keys[r].width)) {
                    vec[k].mean.resize(static_cast<std::size_t>(keys[r].width));
                    vec[k].variance.resize(static_cast<std::size_t>(keys[r].width));
                }
                vec[k].mean[static_cast<std::size_t>(keys[r].comp)] = st.real_mean[r];
                vec[k].variance[static_cast<std::size_t>(keys[r].comp)] = st.real_var[r];
            }
        }
    }

    static bool value_occurs(const std::vector<double> & col, double v)
    {
        for (double x : col) if (x == v) return true;
        return false;
    }

    void load_sidecar(const std::string & path)
    {
        std::ifstream f(path.c_str());
        if (!f.is_open()) return;
        std::string key;
        std::size_t n_particles = 0;
        for (std::string line; std::getline(f, line);) {
            std::istringstream is(line);
            if (!(is >> key)) continue;
            if (key == "n_particles") is >> n_particles;
            else if (key == "ess") is >> ess_;
            else if (key == "log_evidence") is >> log_evidence_;
            else if (key == "max_log_w") is >> max_log_w_;
            else if (key == "real") {
                std::size_t id, k, width;
                real_estimate e;
                is >> id >> k >> width;
                e.mean.resize(width);
                e.variance.resize(width);
                for (auto & m : e.mean) is >> m;
                for (auto & v : e.variance) is >> v;
                auto & vec = real_[id];
                if (vec.size() <= k) vec.resize(k + 1);
                vec[k] = e;
            } else if (key == "int") {
                std::size_t id, k;
                long long lo;
                int bins;
                is >> id >> k >> lo >> bins;
                int_estimate e;
                double best = -1;
                for (int b = 0; b < bins; ++b) {
                    double p;
                    is >> p;
                    if (p > 0) e.distribution[static_cast<int>(lo + b)] = p;
                    if (p > best) { best = p; e.map = static_cast<int>(lo + b); }
                }
                e.num_points = n_particles;
                auto & vec = int_[id];
                if (vec.size() <= k) vec.resize(k + 1);
                vec[k] = e;
            }
        }
    }

    std::map<std::size_t, std::vector<int_estimate>> int_;
    std::map<std::size_t, std::vector<real_estimate>> real_;
    std::vector<std::string> ids_;
    std::string file_name_;
    double ess_ = std::numeric_limits<double>::quiet_NaN();
    double log_evidence_ = std::numeric_limits<double>::quiet_NaN();
    double max_log_w_ = std::numeric_limits<double>::quiet_NaN();
};

}  // end namespace cpprob
#endif  // CPPROB_STATS_PRINTER_HPP
