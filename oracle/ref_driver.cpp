// ORACLE / _ref — TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by or called from the product path.
//
// A thin C driver around the REFERENCE'S OWN post-processing and serialization code, compiled from the sources
// where they lie under /root/reference (recipe: oracle/Makefile, target _ref; output oracle/_ref/):
//     include/cpprob/serialization.hpp                      operator<< / operator>> for pair, vector, tuple (:41-257)
//     include/cpprob/ndarray.hpp                            NDArray<double> text I/O (:273-334)
//     include/cpprob/postprocess/empirical_distribution.hpp EmpiricalDistribution<T> (:16-147)
//     include/cpprob/postprocess/stats_printer.hpp          StatsPrinter (:22-121)
//     include/cpprob/utils.hpp, include/cpprob/traits.hpp   (helpers the above include)
//     include/cpprob/distributions/utils_{discrete,uniform_smallint,poisson}.hpp   logpdf<> of the three distributions no
//                                                           reference test pins (:17-27, :17-27, :17-36)
//     include/cpprob/distributions/utils_normal_distribution.hpp (:20-45), utils_uniform_real.hpp (:21-31),
//     utils_multivariate_normal.hpp (:20-33) + multivariate_normal.hpp   the log-pdfs of the README / BASELINE models
// None of those files is copied or modified.  The third-party headers they name and this image lacks are
// stood in for by oracle/ref_shim/ (Boost type traits, any, filesystem::path/exists, function types, Boost.Random's
// distribution classes, the FlatBuffers runtime and cppzmq as do-nothing classes: see oracle/ref_sis.cpp, which links
// the reference's whole SIS loop against the same stand-ins).
//
// What this pins: the posterior-file grammar as the reference writes and
// parses it, and the StatsPrinter / EmpiricalDistribution arithmetic and console text.  The writer's stream state
// is the one of StateInfer::dump_predicts (src/cpprob/state.cpp:262-267): precision(digits10 = 15), std::scientific;
// values reach the stream through cpprob::any, which forwards to the value's own operator<< (any.hpp:112-117), so a
// vector<pair<size_t, double>> streamed with that state produces the same bytes.
#include <cstring>
#include <fstream>
#include <limits>
#include <numeric>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "cpprob/postprocess/stats_printer.hpp"

// The log-pdf headers of the reference.  Their logpdf<> bodies need nothing but the accessors
// of a Boost.Random distribution type (and pi): compiled unmodified; Boost's classes are stood in for by accessor-only
// shims (oracle/ref_shim/boost/random/, boost/math/constants: the correctly rounded pi), the FlatBuffers and Boost.Math
// names their CSIS members mention by declaration-only stubs.
namespace cpprob { template<class IntType, class RealType> class min_max_discrete_distribution; }   // named by proposal<uniform_smallint>
#include "cpprob/distributions/utils_discrete.hpp"
#include "cpprob/distributions/utils_uniform_smallint.hpp"
#include "cpprob/distributions/utils_poisson.hpp"
#include "cpprob/distributions/utils_normal_distribution.hpp"
#include "cpprob/distributions/utils_uniform_real.hpp"
#include "cpprob/distributions/utils_multivariate_normal.hpp"

namespace {

std::string g_text;

template<class T>
std::string write_record(const std::vector<std::pair<std::size_t, T>> & predicts, double log_w)
{
    std::ostringstream f;
    f.precision(std::numeric_limits<double>::digits10);          // state.cpp:265
    using cpprob::operator<<;
    f << std::scientific << std::make_pair(predicts, log_w) << std::endl;   // state.cpp:266
    return f.str();
}

int put(const std::string & s, char * out, int cap)
{
    if (static_cast<int>(s.size()) + 1 > cap) return -static_cast<int>(s.size()) - 1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return static_cast<int>(s.size());
}

template<class T>
bool parse_record(const char * line, std::pair<std::vector<std::pair<std::size_t, T>>, double> & rec)
{
    std::istringstream iss(line);                                 // stats_printer.hpp:100-101
    using cpprob::operator>>;
    return static_cast<bool>(iss >> rec);
}

}  // namespace

extern "C" {

const char * ref_describe(void)
{
    return "reference code compiled from /root/reference/include: cpprob/serialization.hpp, cpprob/ndarray.hpp, "
           "cpprob/postprocess/empirical_distribution.hpp, cpprob/postprocess/stats_printer.hpp, "
           "cpprob/distributions/utils_discrete.hpp, utils_uniform_smallint.hpp, utils_poisson.hpp, "
           "utils_normal_distribution.hpp, utils_uniform_real.hpp, utils_multivariate_normal.hpp, multivariate_normal.hpp";
}

// ---- writer: serialization.hpp operator<< with dump_predicts' stream state --------------------------------------
int ref_write_real_record(const unsigned long long * ids, const double * vals, int n, double log_w, char * out, int cap)
{
    std::vector<std::pair<std::size_t, double>> p;
    for (int i = 0; i < n; ++i) p.emplace_back(static_cast<std::size_t>(ids[i]), vals[i]);
    return put(write_record(p, log_w), out, cap);
}

int ref_write_int_record(const unsigned long long * ids, const long long * vals, int n, double log_w, char * out, int cap)
{
    // integral predicts are std::size_t / int in the reference's models; both print bare digits
    std::vector<std::pair<std::size_t, long long>> p;
    for (int i = 0; i < n; ++i) p.emplace_back(static_cast<std::size_t>(ids[i]), vals[i]);
    return put(write_record(p, log_w), out, cap);
}

// vector-valued predicts (NDArray<double>, state.hpp:328-340): widths[i] components per predict
int ref_write_ndarray_record(const unsigned long long * ids, const int * widths, const double * vals, int n, double log_w,
                             char * out, int cap)
{
    std::vector<std::pair<std::size_t, cpprob::NDArray<double>>> p;
    const double * v = vals;
    for (int i = 0; i < n; ++i) {
        p.emplace_back(static_cast<std::size_t>(ids[i]), cpprob::NDArray<double>(std::vector<double>(v, v + widths[i])));
        v += widths[i];
    }
    return put(write_record(p, log_w), out, cap);
}

// ---- parser: serialization.hpp operator>> exactly as StatsPrinter::load_distr calls it ---------------------------
// returns the number of predicts, or -1 if the reference's parser rejects the line, or -2 if cap is too small
int ref_parse_real_record(const char * line, unsigned long long * ids, double * vals, int cap, double * log_w)
{
    std::pair<std::vector<std::pair<std::size_t, double>>, double> rec;
    if (!parse_record(line, rec)) return -1;
    if (static_cast<int>(rec.first.size()) > cap) return -2;
    for (std::size_t i = 0; i < rec.first.size(); ++i) { ids[i] = rec.first[i].first; vals[i] = rec.first[i].second; }
    *log_w = rec.second;
    return static_cast<int>(rec.first.size());
}

int ref_parse_int_record(const char * line, unsigned long long * ids, int * vals, int cap, double * log_w)
{
    std::pair<std::vector<std::pair<std::size_t, int>>, double> rec;       // T = int, stats_printer.hpp:83
    if (!parse_record(line, rec)) return -1;
    if (static_cast<int>(rec.first.size()) > cap) return -2;
    for (std::size_t i = 0; i < rec.first.size(); ++i) { ids[i] = rec.first[i].first; vals[i] = rec.first[i].second; }
    *log_w = rec.second;
    return static_cast<int>(rec.first.size());
}

// parse with the reference's parser, print again with the reference's writer: a line the engine wrote must come back
// byte for byte (kind 0 real, 1 int)
int ref_reprint_record(const char * line, int kind, char * out, int cap)
{
    if (kind == 0) {
        std::pair<std::vector<std::pair<std::size_t, double>>, double> rec;
        if (!parse_record(line, rec)) return -1;
        return put(write_record(rec.first, rec.second), out, cap);
    }
    std::pair<std::vector<std::pair<std::size_t, int>>, double> rec;
    if (!parse_record(line, rec)) return -1;
    return put(write_record(rec.first, rec.second), out, cap);
}

// does StatsPrinter's own value type for .real files (NDArray<double>, stats_printer.hpp:84) accept the line?
int ref_parse_real_record_ndarray(const char * line)
{
    std::pair<std::vector<std::pair<std::size_t, cpprob::NDArray<double>>>, double> rec;
    return parse_record(line, rec) ? static_cast<int>(rec.first.size()) : -1;
}

// ---- EmpiricalDistribution -----------------------------------------------------------------------------------------
int ref_empirical_real(const double * x, const double * log_w, unsigned long long n, double * mean, double * variance)
{
    cpprob::EmpiricalDistribution<cpprob::NDArray<double>> d;
    for (unsigned long long i = 0; i < n; ++i) d.add_point(cpprob::NDArray<double>(x[i]), log_w[i]);
    const auto m = d.mean();                                      // stats_printer.hpp:53-55
    const auto v = d.variance(m);
    *mean = static_cast<double>(m);
    *variance = static_cast<double>(v);
    return 0;
}

// values/probs: the std::map<int,double> of distribution() in key order; returns its size (or -needed)
int ref_empirical_int(const int * x, const double * log_w, unsigned long long n, int cap, int * values, double * probs,
                      int * map_value, unsigned long long * num_points)
{
    cpprob::EmpiricalDistribution<int> d;
    for (unsigned long long i = 0; i < n; ++i) d.add_point(x[i], log_w[i]);
    const auto distr = d.distribution();                          // stats_printer.hpp:70-75
    if (static_cast<int>(distr.size()) > cap) return -static_cast<int>(distr.size());
    int k = 0;
    for (const auto & kv : distr) { values[k] = kv.first; probs[k] = kv.second; ++k; }
    *map_value = d.max_a_posteriori(distr);
    *num_points = d.num_points();
    return k;
}

// ---- logpdf<> of utils_normal_distribution.hpp:20-45 (kind 0), utils_uniform_real.hpp:21-31 (1),
// utils_uniform_smallint.hpp:17-27 (2), utils_discrete.hpp:17-27 (3), utils_poisson.hpp:17-36 (4) ------
// kind ids and parameter layout are those of include/cpprob_sis.h (CPPROB_SIS_DIST_*)
int ref_logpdf(int kind, const double * q, int nq, const double * x, unsigned long long n, double * out)
{
    for (unsigned long long i = 0; i < n; ++i) {
        switch (kind) {
        case 0: {
            const boost::random::normal_distribution<double> d(q[0], q[1]);
            out[i] = cpprob::logpdf<boost::random::normal_distribution<double>>()(d, x[i]);
            break;
        }
        case 1: {
            const boost::random::uniform_real_distribution<double> d(q[0], q[1]);
            out[i] = cpprob::logpdf<boost::random::uniform_real_distribution<double>>()(d, x[i]);
            break;
        }
        case 2: {
            const boost::random::uniform_smallint<long long> d(static_cast<long long>(q[0]), static_cast<long long>(q[1]));
            out[i] = cpprob::logpdf<boost::random::uniform_smallint<long long>>()(d, static_cast<long long>(x[i]));
            break;
        }
        case 3: {
            const boost::random::discrete_distribution<long long, double> d(q, q + nq);
            out[i] = cpprob::logpdf<boost::random::discrete_distribution<long long, double>>()(d, static_cast<long long>(x[i]));
            break;
        }
        case 4: {
            const boost::random::poisson_distribution<int, double> d(q[0]);
            out[i] = cpprob::logpdf<boost::random::poisson_distribution<int, double>>()(d, static_cast<int>(x[i]));
            break;
        }
        default: return -1;
        }
    }
    return 0;
}

// ---- logpdf<multivariate_normal_distribution> (utils_multivariate_normal.hpp:20-33): n points of dimension dim, the
// distribution built from (mean[dim], covariance[dim]) as models.hpp does (multivariate_normal.hpp:211-213; the components
// get sigma = sqrt(covariance), :178-186)
int ref_logpdf_mvn(const double * mean, const double * covariance, int dim, const double * x, unsigned long long n, double * out)
{
    const cpprob::multivariate_normal_distribution<double> d(mean, mean + dim, covariance, covariance + dim);
    for (unsigned long long i = 0; i < n; ++i) {
        const cpprob::NDArray<double> xi(std::vector<double>(x + i * dim, x + (i + 1) * dim));
        out[i] = cpprob::logpdf<cpprob::multivariate_normal_distribution<double>>()(d, xi);
    }
    return 0;
}

// ---- StatsPrinter: the console text of `std::cout << cpprob::StatsPrinter{prefix} << std::endl` (src/main.cpp:103-107)
const char * ref_stats_text(const char * prefix)
{
    std::ostringstream os;
    os << cpprob::StatsPrinter{prefix} << std::endl;
    g_text = os.str();
    return g_text.c_str();
}

}  // extern "C"

#ifdef CPPROB_REF_MAIN
// oracle/_ref/ref_stats_printer <prefix>: the reference's StatsPrinter as a process (it calls std::exit on a malformed line)
int main(int argc, char ** argv)
{
    if (argc != 2) { std::cerr << "usage: ref_stats_printer <posterior file prefix>\n"; return 2; }
    std::cout << cpprob::StatsPrinter{argv[1]} << std::endl;
    return 0;
}
#endif
