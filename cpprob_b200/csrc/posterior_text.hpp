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
// what this writer produces, with the text of a trace block formatted by all host cores and written
// in a few large writes instead of three open/append/close per trace.
#ifndef CPPROB_B200_POSTERIOR_TEXT_HPP
#define CPPROB_B200_POSTERIOR_TEXT_HPP

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/statvfs.h>
#include <sys/types.h>
#include <unistd.h>

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
        for (const auto & s : slots) (s.is_int ? int_ids_ : real_ids_).push_back(s);   // program order = row order
    }
    ~posterior_writer()
    {
        if (f_real_ >= 0) ::close(f_real_);
        if (f_int_ >= 0) ::close(f_int_);
    }
    posterior_writer(const posterior_writer &) = delete;
    posterior_writer & operator=(const posterior_writer &) = delete;

    bool open()
    {
        // appended to, never truncated (ios::app in the reference); offsets are tracked here so that the
        // slices of a block can be written concurrently with pwrite
        if (!real_ids_.empty()) {
            f_real_ = ::open((prefix_ + ".real").c_str(), O_RDWR | O_CREAT, 0666);
            if (f_real_ < 0) return false;
            off_real_ = ::lseek(f_real_, 0, SEEK_END);
        }
        if (!int_ids_.empty()) {
            f_int_ = ::open((prefix_ + ".int").c_str(), O_RDWR | O_CREAT, 0666);
            if (f_int_ < 0) return false;
            off_int_ = ::lseek(f_int_, 0, SEEK_END);
        }
        return true;
    }

    // A writer that owns one byte range of files somebody else has already extended to their final size (multi-GPU
    // emission: rank r's records start where rank r-1's end).  Nothing here changes the file size.
    bool open_at(off_t real_offset, off_t int_offset)
    {
        presized_ = true;
        if (!real_ids_.empty()) {
            f_real_ = ::open((prefix_ + ".real").c_str(), O_RDWR | O_CREAT, 0666);
            if (f_real_ < 0) return false;
            off_real_ = real_offset;
        }
        if (!int_ids_.empty()) {
            f_int_ = ::open((prefix_ + ".int").c_str(), O_RDWR | O_CREAT, 0666);
            if (f_int_ < 0) return false;
            off_int_ = int_offset;
        }
        return true;
    }

    // Appends `bytes` of room to <prefix>.<kind> and returns the old end (-1 on error): the offset of the first new byte.
    static off_t extend_file(const std::string & path, unsigned long long bytes)
    {
        const int fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0666);
        if (fd < 0) return -1;
        const off_t end = ::lseek(fd, 0, SEEK_END);
        struct statvfs vfs;
        const bool room = ::fstatvfs(fd, &vfs) != 0 || static_cast<unsigned long long>(vfs.f_bavail) * vfs.f_frsize > bytes + (64ull << 20);
        const bool ok = end >= 0 && room && ::ftruncate(fd, end + static_cast<off_t>(bytes)) == 0;
        if (end >= 0 && !room) errno = ENOSPC;
        ::close(fd);
        return ok ? end : static_cast<off_t>(-1);
    }

    // One trace block -> text.  The block is cut into slices that are formatted concurrently (one buffer
    // per slice) and written in order, so the file content does not depend on the thread count.
    bool append(const cpprob_sis_block & blk)
    {
        if (f_real_ >= 0 && !append_kind(blk, false)) return false;
        if (f_int_ >= 0 && !append_kind(blk, true)) return false;
        return true;
    }

    // Text of a whole batch formatted elsewhere (on the GPU, text_kernels.cuh), appended to the file.
    // Concurrent pwrites to ONE file serialise on the inode lock (measured 3.9 GB/s on tmpfs whatever the thread
    // count), so the file is extended with ftruncate and the text is copied into a shared mapping of the new
    // range by all host cores: page allocation and the copy then scale with the thread count.  Free space is
    // checked first (a full filesystem would otherwise surface as SIGBUS); anything mmap cannot serve falls
    // back to pwrite.
    bool write_text(bool is_int, const char * data, size_t n)
    {
        const int fd = is_int ? f_int_ : f_real_;
        if (fd < 0) return n == 0;
        if (n == 0) return true;
        off_t & file_off = is_int ? off_int_ : off_real_;
        unsigned hw = std::thread::hardware_concurrency();
        if (const char * s = std::getenv("CPPROB_SIS_WRITER_THREADS")) hw = static_cast<unsigned>(std::atoi(s));
        const size_t kPiece = 4u << 20;
        const size_t n_threads = std::max<size_t>(1, std::min<size_t>({static_cast<size_t>(hw ? hw : 1), (n + kPiece - 1) / kPiece, 64}));
        const char * mode = std::getenv("CPPROB_SIS_FILE_IO");
        const bool want_mmap = !(mode && std::strcmp(mode, "pwrite") == 0);
        char * map = nullptr;
        size_t lead = 0;
        if (want_mmap) {
            struct statvfs vfs;
            const bool room = ::fstatvfs(fd, &vfs) != 0 || static_cast<unsigned long long>(vfs.f_bavail) * vfs.f_frsize > n + (64ull << 20);
            if (!room) { errno = ENOSPC; return false; }
            const off_t page = static_cast<off_t>(::sysconf(_SC_PAGESIZE));
            const off_t map_off = file_off / page * page;
            lead = static_cast<size_t>(file_off - map_off);
            if (presized_ || ::ftruncate(fd, file_off + static_cast<off_t>(n)) == 0) {
                void * m = ::mmap(nullptr, lead + n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, map_off);
                if (m != MAP_FAILED) map = static_cast<char *>(m);
            }
        }
        const size_t per = ((n + n_threads - 1) / n_threads + 4095) / 4096 * 4096;
        std::vector<char> ok(n_threads, 1);
        auto piece = [&](size_t t) {
            const size_t lo = t * per, hi = std::min(n, lo + per);
            if (lo >= hi) return;
            if (map) std::memcpy(map + lead + lo, data + lo, hi - lo);
            else ok[t] = write_all(fd, data + lo, hi - lo, file_off + static_cast<off_t>(lo)) ? 1 : 0;
        };
        std::vector<std::thread> pool;
        for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(piece, t);
        piece(0);
        for (auto & th : pool) th.join();
        if (map) ::munmap(map, lead + n);
        file_off += static_cast<off_t>(n);
        for (char c : ok) if (!c) return false;
        return true;
    }

    // the line of record i of `blk`, formatted on the host (reference path; also the fallback for records the
    // GPU formatter reports as ambiguous)
    std::string format_record(const cpprob_sis_block & blk, bool is_int, size_t i) const
    {
        text_buffer b;
        format_slice(blk, is_int, i, i + 1, b);
        return std::string(b.p.get(), b.len);
    }

    bool has_real() const { return f_real_ >= 0; }
    bool has_int() const { return f_int_ >= 0; }
    const std::vector<cpprob_sis_slot> & slots(bool is_int) const { return is_int ? int_ids_ : real_ids_; }

    // finish_infer: .ids, and removal of the kinds that had no predicts
    bool finish(const std::vector<std::string> & ids)
    {
        bool ok = true;
        if (f_real_ >= 0) { ok = ::close(f_real_) == 0 && ok; f_real_ = -1; }
        if (f_int_ >= 0) { ok = ::close(f_int_) == 0 && ok; f_int_ = -1; }
        ok = write_ids(prefix_, ids) && ok;
        if (int_ids_.empty()) std::remove((prefix_ + ".int").c_str());
        if (real_ids_.empty()) std::remove((prefix_ + ".real").c_str());
        std::remove((prefix_ + ".any").c_str());   // no any-typed predicts exist on the device path
        return ok;
    }

private:
    // grow-only, never zero-filled text buffer of one formatting thread
    struct text_buffer {
        std::unique_ptr<char[]> p;
        size_t cap = 0, len = 0;
        char * reserve(size_t n)
        {
            if (n > cap) {
                p.reset(new char[n]);
                cap = n;
            }
            return p.get();
        }
    };

    void format_slice(const cpprob_sis_block & blk, bool is_int, size_t i0, size_t i1, text_buffer & out) const
    {
        const std::vector<cpprob_sis_slot> & ids = is_int ? int_ids_ : real_ids_;
        // "(" "[" n*( "(" id " " value ")" " " ) "]" " " logw ")" "\n"; a vector value is "[v0 v1 ...]"
        size_t per_line = 8 + 32;
        for (const auto & s : ids) per_line += 24 + static_cast<size_t>(s.width) * 26;
        char * p = out.reserve(per_line * (i1 - i0));
        char * const begin = p;
        for (size_t i = i0; i < i1; ++i) {
            *p++ = '(';
            *p++ = '[';
            for (size_t r = 0; r < ids.size(); ++r) {
                if (r) *p++ = ' ';
                *p++ = '(';
                p = format_int(p, ids[r].id);
                *p++ = ' ';
                const size_t row = static_cast<size_t>(ids[r].row);
                if (is_int) {
                    p = format_int(p, blk.int_rows[row * blk.stride + i]);
                } else if (ids[r].width == 1) {
                    p = format_double(p, blk.real_rows[row * blk.stride + i]);
                } else {                                   // NDArray vector: ndarray.hpp:273-288 -> "[v0 v1 ...]"
                    *p++ = '[';
                    for (int c = 0; c < ids[r].width; ++c) {
                        if (c) *p++ = ' ';
                        p = format_double(p, blk.real_rows[(row + c) * blk.stride + i]);
                    }
                    *p++ = ']';
                }
                *p++ = ')';
            }
            *p++ = ']';
            *p++ = ' ';
            p = format_double(p, blk.log_w[i]);
            *p++ = ')';
            *p++ = '\n';
        }
        out.len = static_cast<size_t>(p - begin);
    }

    static bool write_all(int fd, const char * p, size_t n, off_t off)
    {
        while (n > 0) {
            const ssize_t w = ::pwrite(fd, p, n, off);
            if (w < 0) return false;
            p += w;
            n -= static_cast<size_t>(w);
            off += w;
        }
        return true;
    }

    bool append_kind(const cpprob_sis_block & blk, bool is_int)
    {
        const int fd = is_int ? f_int_ : f_real_;
        off_t & file_off = is_int ? off_int_ : off_real_;
        const size_t kSlice = 1 << 16;                      // records per slice
        const size_t n_slices = (blk.n + kSlice - 1) / kSlice;
        unsigned hw = std::thread::hardware_concurrency();
        if (const char * s = std::getenv("CPPROB_SIS_WRITER_THREADS")) hw = static_cast<unsigned>(std::atoi(s));
        const size_t n_threads = std::max<size_t>(1, std::min<size_t>({static_cast<size_t>(hw ? hw : 1), n_slices, 32}));
        // waves of n_threads slices: format concurrently, then (offsets known) write concurrently
        if (bufs_.size() < n_threads) bufs_.resize(n_threads);
        std::vector<text_buffer> & bufs = bufs_;
        std::vector<off_t> offs(n_threads);
        std::vector<char> ok(n_threads, 1);
        auto run_wave = [&](size_t wave, const std::function<void(size_t)> & fn) {
            if (wave == 1) { fn(0); return; }
            std::vector<std::thread> pool;
            pool.reserve(wave);
            for (size_t t = 0; t < wave; ++t) pool.emplace_back(fn, t);
            for (auto & th : pool) th.join();
        };
        for (size_t s0 = 0; s0 < n_slices; s0 += n_threads) {
            const size_t wave = std::min(n_threads, n_slices - s0);
            run_wave(wave, [&](size_t t) {
                format_slice(blk, is_int, (s0 + t) * kSlice, std::min<size_t>(blk.n, (s0 + t + 1) * kSlice), bufs[t]);
            });
            for (size_t t = 0; t < wave; ++t) {
                offs[t] = file_off;
                file_off += static_cast<off_t>(bufs[t].len);
            }
            run_wave(wave, [&](size_t t) { ok[t] = write_all(fd, bufs[t].p.get(), bufs[t].len, offs[t]) ? 1 : 0; });
            for (size_t t = 0; t < wave; ++t) if (!ok[t]) return false;
        }
        return true;
    }

    std::string prefix_;
    std::vector<cpprob_sis_slot> real_ids_, int_ids_;   // the real / int predict slots in program order
    int f_real_ = -1, f_int_ = -1;
    std::vector<text_buffer> bufs_;
    off_t off_real_ = 0, off_int_ = 0;
    bool presized_ = false;
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
            std::fprintf(f, "real %d %d %d", s.id, s.k, s.width);
            for (int c = 0; c < s.width; ++c) std::fprintf(f, " %.17g", st.real_mean[s.row + c]);
            for (int c = 0; c < s.width; ++c) std::fprintf(f, " %.17g", st.real_var[s.row + c]);
            std::fprintf(f, "\n");
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
