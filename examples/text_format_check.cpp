// CPU check of the device-side %.15e formatter (cpprob_b200/csrc/text_format.cuh is __host__ __device__):
// byte-for-byte against printf("%.15e") over random bit patterns, random magnitudes, integers, powers of
// two and ten, exact ties, subnormals and the specials.  usage: text_format_check [n_random]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "text_format.cuh"

static long long checked = 0, mismatches = 0, ambiguous_count = 0;

static void check(double v)
{
    char ours[40], ref[40];
    bool amb = false;
    char * e = cpprob::text::format_e15(ours, v, &amb);
    *e = 0;
    std::snprintf(ref, sizeof ref, "%.15e", v);
    ++checked;
    if (amb) { ++ambiguous_count; return; }          // the engine lets the host format these
    if (std::strcmp(ours, ref) != 0) {
        if (mismatches < 20) std::printf("MISMATCH %a: ours %s ref %s\n", v, ours, ref);
        ++mismatches;
    }
}

int main(int argc, char ** argv)
{
    const long long n = argc > 1 ? std::atoll(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    for (long long i = 0; i < n; ++i) {               // random bit patterns: every exponent, incl. subnormals / nan / inf
        const std::uint64_t b = rng();
        double v;
        std::memcpy(&v, &b, sizeof v);
        check(v);
    }
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::normal_distribution<double> g(0.0, 1.0);
    for (long long i = 0; i < n; ++i) {               // the magnitudes the engine actually prints
        check(g(rng));
        check(-std::fabs(g(rng)) * 10);
        check(u(rng) * std::pow(10.0, static_cast<int>(u(rng) * 40) - 20));
    }
    for (long long i = -100000; i <= 100000; ++i) check(static_cast<double>(i));
    for (int i = 0; i < 200000; ++i) {                // integers and half-integers around 2^52..2^53: exact ties at digit 16
        const double base = 4503599627370496.0 + static_cast<double>(rng() % 4503599627370496ull);
        check(base);
        check(base + 0.5);
        check(5000000000000000.5 + i);
        check(1000000000000000.5 + i * 7.0);
    }
    for (int e = -1074; e <= 1023; ++e) {
        check(std::ldexp(1.0, e));
        check(std::ldexp(3.0, e));
        check(std::nextafter(std::ldexp(1.0, e), 0.0));
        check(std::nextafter(std::ldexp(1.0, e), INFINITY));
    }
    for (int k = -323; k <= 308; ++k) {
        const double p = std::pow(10.0, k);
        check(p);
        check(std::nextafter(p, 0.0));
        check(std::nextafter(p, INFINITY));
        check(9.999999999999999 * p);
        check(9.9999999999999995 * p);
    }
    const double specials[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                               0.1, 0.2, 0.3, 0.5, 1.5, 2.5, 1e15, 1e16, 1e22, 1e23, 123456789012345678.0, 2.32353, 1.05882};
    for (double s : specials) check(s);
    std::printf("checked %lld values: %lld mismatches, %lld ambiguous\n", checked, mismatches, ambiguous_count);
    return mismatches ? 1 : 0;
}
