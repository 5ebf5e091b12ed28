// cpprob-b200: the posterior files, byte-compatible with the reference's writer.
//
// Reference behaviour restated here (/root/reference):
//   * StateInfer::finish_trace (src/cpprob/state.cpp:193-202) appends, per trace, one line to each
//     of <out>.int / <out>.real / <out>.any; dump_predicts (:262-267) opens with ios::app, sets
//     precision(digits10 = 15) and std::scientific, and streams
//     std::make_pair(vector<pair<size_t, any>>, log_w), which the grammar of
//     include/cpprob/serialization.hpp:41-46 (pair -> "(a b)") and :71-98 (vector -> "[a b c]")
//     turns into        ([(id v) (id v) ...] logw)\n        — an empty list prints "([] logw)".
//     Integral values print bare, doubles as %.15e (any.hpp:112-117 streams the held value).
//   * finish_infer (:164-180) writes <out>.ids (dump_ids :250-260: one address per line, line index
//     = id, file truncated) and removes every kind whose list was empty for all traces.
// The net effect for a model with a fixed predict structure is: a kind that has predicts gets one
// line per trace appended; a kind that has none ends up removed (even if it pre-existed).  That is
// what this writer produces, one buffered write per trace block instead of three open/append/close
// per trace.
#ifndef CPPROB_B200_POSTERIOR_TEXT_HPP
#define CPPROB_B200_POSTERIOR_TEXT_HPP

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "cpprob_sis.h"

namespace cpprob {
namespace text {

// %.15e, exactly as `os << std::scientific << std::setprecision(15) << v`
inline char * format_double(char * p, double v)
{
    if (std::isnan(v)) {
        const char * s = std::signbit(v) ? "-nan" : "nan";
        const size_t n = std::strlen(s);
        std::memcpy(p, s, n);
        return p + n;
    }
    if (std::isinf(v)) {
        const char * s = v < 0 ? "-inf" : "inf";
        const size_t n = std::strlen(s);
        std::memcpy(p, s, n);
        return p + n;
    }
    const auto r = std::to_chars(p, p + 32, v, std::chars_format::scientific, 15);
    return r.ptr;
}

inline char * format_int(char * p, long long v)
{
    const auto r = std::to_chars(p, p + 24, v);
    return r.ptr;
}

// dump_ids (state.cpp:250-260): one address per line, line index = id; the file is truncated
inline bool write_ids(const std::string & prefix, const std::vector<std::string> & ids)
{
    std::FILE * f = std::fopen((prefix + ".ids").c_str(), "wb");
    if (!f) return false;
    for (const auto & s : ids) {
        std::fputs(s.c_str(), f);
        std::fputc('\n', f);
    }
    return std::fclose(f) == 0;
}

class posterior_writer {
public:
    posterior_writer(const std::string & prefix, const std::vector<cpprob_sis_slot> & slots) : prefix_(prefix)
    {
        for (const auto & s : slots) {
            if (s.is_int) {
                if (int_ids_.size() <= static_cast<size_t>(s.row)) int_ids_.resize(static_cast<size_t>(s.row) + 1);
                int_ids_[static_cast<size_t>(s.row)] = s.id;
            } else {
                if (real_ids_.size() <= static_cast<size_t>(s.row)) real_ids_.resize(static_cast<size_t>(s.row) + 1);
                real_ids_[static_cast<size_t>(s.row)] = s.id;
            }
        }
    }
    ~posterior_writer()
    {
        if (f_real_) std::fclose(f_real_);
        if (f_int_) std::fclose(f_int_);
    }
    posterior_writer(const posterior_writer &) = delete;
    posterior_writer & operator=(const posterior_writer &) = delete;

    bool open()
    {
        if (!real_ids_.empty()) {
            f_real_ = std::fopen((prefix_ + ".real").c_str(), "ab");
            if (!f_real_) return false;
        }
        if (!int_ids_.empty()) {
            f_int_ = std::fopen((prefix_ + ".int").c_str(), "ab");
            if (!f_int_) return false;
        }
        return true;
    }

    bool append(const cpprob_sis_block & blk)
    {
        if (f_real_) {
            // "(" "[" n*( "(" id " " 23 ")" " " ) "]" " " 24 ")" "\n"
            const size_t per_line = 8 + real_ids_.size() * 48 + 32;
            buf_.resize(per_line * 4096);
            size_t done = 0;
            while (done < blk.n) {
                const size_t n = std::min<size_t>(4096, blk.n - done);
                char * p = buf_.data();
                for (size_t i = done; i < done + n; ++i) {
                    *p++ = '(';
                    *p++ = '[';
                    for (size_t r = 0; r < real_ids_.size(); ++r) {
                        if (r) *p++ = ' ';
                        *p++ = '(';
                        p = format_int(p, real_ids_[r]);
                        *p++ = ' ';
                        p = format_double(p, blk.real_rows[r * blk.stride + i]);
                        *p++ = ')';
                    }
                    *p++ = ']';
                    *p++ = ' ';
                    p = format_double(p, blk.log_w[i]);
                    *p++ = ')';
                    *p++ = '\n';
                }
                if (std::fwrite(buf_.data(), 1, static_cast<size_t>(p - buf_.data()), f_real_) != static_cast<size_t>(p - buf_.data())) return false;
                done += n;
            }
        }
        if (f_int_) {
            const size_t per_line = 8 + int_ids_.size() * 40 + 32;
            buf_.resize(per_line * 4096);
            size_t done = 0;
            while (done < blk.n) {
                const size_t n = std::min<size_t>(4096, blk.n - done);
                char * p = buf_.data();
                for (size_t i = done; i < done + n; ++i) {
                    *p++ = '(';
                    *p++ = '[';
                    for (size_t r = 0; r < int_ids_.size(); ++r) {
                        if (r) *p++ = ' ';
                        *p++ = '(';
                        p = format_int(p, int_ids_[r]);
                        *p++ = ' ';
                        p = format_int(p, blk.int_rows[r * blk.stride + i]);
                        *p++ = ')';
                    }
                    *p++ = ']';
                    *p++ = ' ';
                    p = format_double(p, blk.log_w[i]);
                    *p++ = ')';
                    *p++ = '\n';
                }
                if (std::fwrite(buf_.data(), 1, static_cast<size_t>(p - buf_.data()), f_int_) != static_cast<size_t>(p - buf_.data())) return false;
                done += n;
            }
        }
        return true;
    }

    // finish_infer: .ids, and removal of the kinds that had no predicts
    bool finish(const std::vector<std::string> & ids)
    {
        bool ok = true;
        if (f_real_) { ok = std::fclose(f_real_) == 0 && ok; f_real_ = nullptr; }
        if (f_int_) { ok = std::fclose(f_int_) == 0 && ok; f_int_ = nullptr; }
        ok = write_ids(prefix_, ids) && ok;
        if (int_ids_.empty()) std::remove((prefix_ + ".int").c_str());
        if (real_ids_.empty()) std::remove((prefix_ + ".real").c_str());
        std::remove((prefix_ + ".any").c_str());   // no any-typed predicts exist on the device path
        return ok;
    }

private:
    std::string prefix_;
    std::vector<int> real_ids_, int_ids_;   // address id of each real / int row
    std::FILE * f_real_ = nullptr;
    std::FILE * f_int_ = nullptr;
    std::vector<char> buf_;
};

// <prefix>.stats: the on-device estimators of the LAST run (the record files may hold older runs
// too, since they are appended to).  Line-oriented `key value...`; read by StatsPrinter's fast path.
inline bool write_stats_sidecar(const std::string & prefix, const cpprob_sis_stats & st,
                                const std::vector<cpprob_sis_slot> & slots, const std::vector<std::string> & ids)
{
    std::FILE * f = std::fopen((prefix + ".stats").c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "cpprob_sis_stats 1\n");
    std::fprintf(f, "n_particles %llu\n", static_cast<unsigned long long>(st.n_particles));
    std::fprintf(f, "n_neg_inf %llu\n", static_cast<unsigned long long>(st.n_neg_inf));
    std::fprintf(f, "m_ref %.17g\n", st.m_ref);
    std::fprintf(f, "max_log_w %.17g\n", st.max_log_w);
    std::fprintf(f, "log_sum_exp %.17g\n", st.log_sum_exp);
    std::fprintf(f, "log_evidence %.17g\n", st.log_evidence);
    std::fprintf(f, "ess %.17g\n", st.ess);
    std::fprintf(f, "n_ids %zu\n", ids.size());
    for (const auto & s : ids) std::fprintf(f, "id %s\n", s.c_str());
    for (const auto & s : slots) {
        if (!s.is_int) {
            std::fprintf(f, "real %d %d %.17g %.17g\n", s.id, s.k, st.real_mean[s.row], st.real_var[s.row]);
        }
    }
    for (const auto & s : slots) {
        if (s.is_int) {
            std::fprintf(f, "int %d %d %lld %d", s.id, s.k, st.int_lo, st.int_bins);
            for (int b = 0; b < st.int_bins; ++b) std::fprintf(f, " %.17g", st.int_prob[static_cast<size_t>(s.row) * st.int_bins + b]);
            std::fprintf(f, "\n");
        }
    }
    return std::fclose(f) == 0;
}

}  // namespace text
}  // namespace cpprob
#endif  // CPPROB_B200_POSTERIOR_TEXT_HPP
