// Host-side check of limits that must fail loudly instead of changing the model silently (headers only, no engine).
#include <cstdio>
#include <stdexcept>
#include <vector>
#include "cpprob/distributions/distributions.hpp"

int main()
{
    int failures = 0;
    // more weights than the inline capacity of discrete_distribution: Boost's has no such limit, so never truncate
    const std::vector<double> nine(9, 1.0);
    try {
        cpprob::discrete_distribution<int, double, 8> d(nine.begin(), nine.end());
        std::printf("FAIL: 9 weights were accepted by a capacity-8 discrete_distribution (max %d)\n", d.max());
        ++failures;
    } catch (const std::length_error & e) {
        std::printf("ok: %s\n", e.what());
    }
    // a larger capacity named at the call site takes them
    cpprob::discrete_distribution<int, double, 16> d16(nine.begin(), nine.end());
    if (d16.max() != 8 || d16.probabilities()[3] != 1.0 / 9.0) { std::printf("FAIL: capacity 16\n"); ++failures; }
    else std::printf("ok: capacity 16 holds 9 weights\n");
    // exactly the capacity is fine
    const std::vector<double> eight(8, 0.5);
    cpprob::discrete_distribution<int, double, 8> d8(eight.begin(), eight.end());
    if (d8.max() != 7) { std::printf("FAIL: capacity 8\n"); ++failures; }
    return failures;
}
