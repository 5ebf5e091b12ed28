// Host twin of the device normal sampler (include/cpprob/random/philox.hpp is __host__ __device__): writes n
// standard normals of the streams first, first+1, ... (one draw each, as cpprob_sis_sample does) to stdout as raw
// little-endian doubles.  tests/test_host_api.py checks their distribution on the CPU; tests/test_device_layer_gpu.py
// checks that the GPU produces the same bits.   usage: zig_check <seed> <first> <n> [draws_per_stream]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cpprob/random/philox.hpp"

int main(int argc, char ** argv)
{
    if (argc < 4) return 2;
    const std::uint64_t seed = std::strtoull(argv[1], nullptr, 0), first = std::strtoull(argv[2], nullptr, 0), n = std::strtoull(argv[3], nullptr, 0);
    const std::uint64_t per = argc > 4 ? std::strtoull(argv[4], nullptr, 0) : 1;
    const cpprob::philox_keys keys(seed);
    std::vector<double> out;
    out.reserve(1 << 16);
    for (std::uint64_t i = 0; i < n; ++i) {
        cpprob::philox_stream rng(keys, first + i);
        for (std::uint64_t d = 0; d < per; ++d) {
            out.push_back(rng.next_std_normal());
            if (out.size() == (1u << 16)) { std::fwrite(out.data(), sizeof(double), out.size(), stdout); out.clear(); }
        }
    }
    std::fwrite(out.data(), sizeof(double), out.size(), stdout);
    return 0;
}
