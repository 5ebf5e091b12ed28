// ORACLE — TEST INFRASTRUCTURE (nothing under cpprob_b200/, include/ or examples/ includes, links or runs this).
//
// Driver of the REFERENCE'S OWN sequential importance sampling: `cpprob::inference(cpprob::StateType::sis, &model, observes,
// n, file)` of /root/reference include/cpprob/cpprob.hpp:173-203, with the reference's own state machinery linked in
// unmodified — src/cpprob/{state,trace,sample,utils,serialization,cpprob,socket}.cpp and its models
// (include/models/models.hpp, src/models/gaussian.cpp = the README program's model, src/models/models.cpp = all_distr).
// oracle/Makefile (target _ref) compiles those files where they lie under /root/reference; nothing of them is copied.
// What this image lacks is third-party only and is stood in for by oracle/ref_shim/: Boost (type traits, any, function
// types, filesystem::path, math::normal, and Boost.Random's distribution classes as accessors + a draw), the FlatBuffers
// runtime and cppzmq (names only: the wire protocol and the sockets belong to the compile / CSIS modes, never entered here).
//
// Two uses:
//   * --replay <file>: the stand-in distributions return the values of <file> (raw float64, program order, trace after
//     trace) instead of drawing.  The reference's own code then computes every log-pdf, accumulates the log-weights, routes
//     the predicts, numbers the addresses and writes <prefix>.real / .int / .ids — for the very values the CUDA path (or the
//     restated oracle) sampled.  tests/test_ref_sis*.py compare those files byte for byte.
//   * no --replay: draws come from the C++ standard library's distributions through the reference's get_rng()
//     (std::mt19937 seeded from random_device, src/cpprob/utils.cpp:16-20): the timed CPU baseline of bench.py
//     (`cpu_baseline.kind = "reference"`).
//
// usage: ref_sis <model> <n_particles> <prefix> <replay file | -> <observation>...
//        prints "seconds <wall time of cpprob::inference>" on stderr; the reference's own progress lines go to stdout.
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <tuple>
#include <vector>

#include "models/models.hpp"
#include "models/gaussian.hpp"
#include "models/poly_adjustment.hpp"
#include "cpprob/cpprob.hpp"

namespace models { void all_distr(int, int); }   // src/models/models.cpp:13 (no header declares it)

namespace {

template<class F, class Obs>
double run(const F & f, const Obs & observes, std::size_t n, const std::string & prefix)
{
    const auto t0 = std::chrono::steady_clock::now();
    cpprob::inference(cpprob::StateType::sis, f, observes, n, prefix);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

template<std::size_t N>
std::array<double, N> as_array(const std::vector<double> & v)
{
    std::array<double, N> a;
    for (std::size_t i = 0; i < N; ++i) a[i] = v[i];
    return a;
}

}  // namespace

int main(int argc, char ** argv)
{
    if (argc < 6) {
        std::fprintf(stderr, "usage: ref_sis <model> <n_particles> <prefix> <replay file | -> <observation>...\n");
        return 2;
    }
    const std::string model = argv[1], prefix = argv[3], replay_file = argv[4];
    const std::size_t n = static_cast<std::size_t>(std::strtoull(argv[2], nullptr, 10));
    std::vector<double> obs;
    for (int i = 5; i < argc; ++i) obs.push_back(std::strtod(argv[i], nullptr));

    std::vector<double> replay;
    if (replay_file != "-") {
        std::ifstream in(replay_file, std::ios::binary | std::ios::ate);
        if (!in) { std::fprintf(stderr, "cannot open %s\n", replay_file.c_str()); return 2; }
        replay.resize(static_cast<std::size_t>(in.tellg()) / sizeof(double));
        in.seekg(0);
        in.read(reinterpret_cast<char *>(replay.data()), static_cast<std::streamsize>(replay.size() * sizeof(double)));
        cpprob_ref_shim::replay().values = replay.data();
        cpprob_ref_shim::replay().n = replay.size();
        cpprob_ref_shim::replay().pos = 0;
    }

    double seconds = -1.0;
    const std::size_t k = obs.size();
    if (model == "gaussian_unknown_mean" && k == 2) {                      // README.md:102-116, src/models/gaussian.cpp
        seconds = run(static_cast<void (*)(double, double)>(&models::gaussian_unknown_mean), std::make_tuple(obs[0], obs[1]), n, prefix);
    } else if (model == "gaussian_unknown_mean_mu" && k == 2) {            // models.hpp:22-35
        seconds = run(&models::gaussian_unknown_mean<double>, std::make_tuple(obs[0], obs[1]), n, prefix);
    } else if (model == "normal_rejection_sampling" && k == 2) {           // models.hpp:82-112
        seconds = run(&models::normal_rejection_sampling<double>, std::make_tuple(obs[0], obs[1]), n, prefix);
    } else if (model == "gaussian_2d_unk_mean") {                          // models.hpp:38-49
        seconds = run(&models::gaussian_2d_unk_mean<double>, std::make_tuple(obs), n, prefix);
    } else if (model == "all_distr") {                                      // src/models/models.cpp:13-47
        seconds = run(&models::all_distr, std::make_tuple(0, 0), n, prefix);
    } else if (model.rfind("poly_adjustment_", 0) == 0 && k == 12) {       // poly_adjustment.hpp:85-95, six (x, y) points
        std::array<std::array<double, 2>, 6> pts;
        for (std::size_t i = 0; i < 6; ++i) pts[i] = {{obs[2 * i], obs[2 * i + 1]}};
        const auto o = std::make_tuple(pts);
        if (model == "poly_adjustment_1") seconds = run(&models::poly_adjustment<1, 6>, o, n, prefix);
        else if (model == "poly_adjustment_2") seconds = run(&models::poly_adjustment<2, 6>, o, n, prefix);
        else if (model == "poly_adjustment_3") seconds = run(&models::poly_adjustment<3, 6>, o, n, prefix);
    } else if (model == "linear_regression" && k % 2 == 0) {               // poly_adjustment.hpp:57-82 (main.cpp's "dyn_linear_reg"; the Builder argument is default-made by call_f_tuple)
        std::vector<std::pair<double, double>> pts;
        for (std::size_t i = 0; i + 1 < k; i += 2) pts.emplace_back(obs[i], obs[i + 1]);
        seconds = run(&models::linear_regression<double>, std::make_tuple(pts), n, prefix);
    } else if (model == "linear_gaussian_1d" && k == 5) {                  // models.hpp:67-80
        seconds = run(&models::linear_gaussian_1d<5>, std::make_tuple(as_array<5>(obs)), n, prefix);
    } else if (model == "linear_gaussian_1d" && k == 8) {
        seconds = run(&models::linear_gaussian_1d<8>, std::make_tuple(as_array<8>(obs)), n, prefix);
    } else if (model == "linear_gaussian_1d" && k == 32) {
        seconds = run(&models::linear_gaussian_1d<32>, std::make_tuple(as_array<32>(obs)), n, prefix);
    } else if (model == "hmm" && k == 9) {                                  // models.hpp:114-141
        seconds = run(&models::hmm<9>, std::make_tuple(as_array<9>(obs)), n, prefix);
    } else if (model == "hmm" && k == 12) {
        seconds = run(&models::hmm<12>, std::make_tuple(as_array<12>(obs)), n, prefix);
    } else if (model == "hmm" && k == 64) {
        seconds = run(&models::hmm<64>, std::make_tuple(as_array<64>(obs)), n, prefix);
    } else if (model == "hmm" && k == 1000) {
        seconds = run(&models::hmm<1000>, std::make_tuple(as_array<1000>(obs)), n, prefix);
    } else {
        std::fprintf(stderr, "ref_sis: no instantiation for model %s with %zu observations\n", model.c_str(), k);
        return 2;
    }
    if (cpprob_ref_shim::replay().active() && cpprob_ref_shim::replay().pos != cpprob_ref_shim::replay().n) {
        std::fprintf(stderr, "ref_sis: %zu of %zu replayed values were consumed\n", cpprob_ref_shim::replay().pos, cpprob_ref_shim::replay().n);
        return 3;
    }
    std::fprintf(stderr, "seconds %.6f\n", seconds);
    return 0;
}
