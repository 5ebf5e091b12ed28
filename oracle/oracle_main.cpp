// ORACLE — TEST INFRASTRUCTURE ONLY (see cpprob_oracle.hpp).
// CLI used to time the restated reference CPU SIS:  oracle_sis <model> <n> <prefix> <faithful|fast> <obs...>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "cpprob_oracle.hpp"

extern "C" double oracle_run(const char *, const double *, int, unsigned long long, const char *, int, unsigned, int);
extern "C" const char * oracle_stats_text(const char *);

int main(int argc, char ** argv)
{
    if (argc < 6) {
        std::cerr << "usage: oracle_sis <model> <n> <prefix> <faithful|fast> <obs...>\n";
        return 2;
    }
    const std::string how = argv[4];
    std::vector<double> obs;
    for (int i = 5; i < argc; ++i) obs.push_back(std::atof(argv[i]));
    const double s = oracle_run(argv[1], obs.data(), static_cast<int>(obs.size()), std::strtoull(argv[2], nullptr, 10), argv[3],
                                how == "fast" ? 1 : 0, 20240607u, 2);
    if (s < 0) {
        std::cerr << "unknown model\n";
        return 1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    // ORACLE_STATS_PRINTER=<path of oracle/_ref/ref_stats_printer>: post-process with the REFERENCE'S OWN StatsPrinter
    // (compiled unmodified from /root/reference) instead of the restated one
    if (const char * sp = std::getenv("ORACLE_STATS_PRINTER")) {
        const std::string cmd = std::string(sp) + " '" + argv[3] + "'";
        if (std::system(cmd.c_str()) != 0) {
            std::cerr << "reference StatsPrinter failed\n";
            return 1;
        }
    } else {
        std::cout << oracle_stats_text(argv[3]);
    }
    const double s2 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cerr << "inference_s " << s << " stats_printer_s " << s2 << "\n";
    return 0;
}
