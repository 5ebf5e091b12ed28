// Repeated cpprob::inference calls from one process: after the first call the engine (CUDA context, streams, device
// tables, buffers) is reused (sis::cached_engine), so a 10,000-particle README inference costs its kernels, one small
// device->host copy and the file append.  Prints the wall time of every call.
// usage: repeat_inference <output prefix> [calls] [particles]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <tuple>
#include "models/gaussian.hpp"
#include "cpprob/cpprob.hpp"
#include "cpprob/postprocess/stats_printer.hpp"

int main(int argc, char ** argv)
{
    const std::string outfile = argc > 1 ? argv[1] : "posterior_sis";
    const int calls = argc > 2 ? std::atoi(argv[2]) : 10;
    const std::size_t samples = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 10'000;
    const auto observes = std::make_tuple(3., 4.);
    for (int i = 0; i < calls; ++i) {
        for (const char * ext : {".real", ".ids", ".stats"}) std::remove((outfile + ext).c_str());
        const auto t0 = std::chrono::steady_clock::now();
        cpprob::inference(cpprob::StateType::sis, &models::gaussian_unknown_mean, observes, samples, outfile);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::printf("call %d: %.3f ms\n", i, ms);
    }
    std::cout << cpprob::StatsPrinter{outfile} << std::endl;
    return 0;
}
