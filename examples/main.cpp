// cpprob-b200: command-line driver for the SIS path, after /root/reference src/main.cpp:116-239
// (`./main --sis --estimate --model unk_mean -n 10000 -o "3 4"`).  Same option names and defaults for
// everything that concerns SIS; the compile / csis / dryrun switches belong to inference compilation,
// which this engine leaves to the reference, and are rejected with a message.  Boost.ProgramOptions /
// Boost.Filesystem are replaced by a few lines of std.
#include <array>
#include <cstdlib>
#include <iostream>
#include <string>
#include <sys/stat.h>
#include <tuple>

#include "cpprob/cpprob.hpp"
#include "cpprob/postprocess/stats_printer.hpp"
#include "cpprob/serialization.hpp"
#include "models/models.hpp"

namespace {

const char * const kModelNames = "{unk_mean,unk_mean_rejection,linear_gaussian,hmm,linear_regression,dyn_linear_reg,unk_mean_2d}";

struct options {
    bool sis = false, estimate = false;
    std::string model, model_folder, observes, observes_file, generated_file = "post";
    std::size_t n_samples = 10000;
};

void usage()
{
    std::cout << "CPProb (B200 SIS engine) options:\n"
                 "  -h [ --help ]                 Print help message\n"
                 "  --sis                         Sequential Importance Sampling: Priors as proposals.\n"
                 "  --estimate                    Estimators.\n"
                 "  --model {unk_mean,unk_mean_rejection,linear_gaussian,hmm,linear_regression,dyn_linear_reg,unk_mean_2d}\n"
                 "                                (SIS) Select the model to be executed\n"
                 "  --model_folder arg            Folder to save the model data. Default: the model name\n"
                 "  -n [ --n_samples ] arg (=10000)  (SIS) Number of particles to be sampled from the posterior.\n"
                 "  -o [ --observes ] arg         (SIS) Values to observe.\n"
                 "  -f [ --observes_file ] arg    (SIS) File with the observed values.\n"
                 "  --generated_file arg (=post)  (SIS | Estimate) File for the samples from the posterior.\n";
}

// the observation tuple of a model function: void(A...) -> std::tuple<decay_t<A>...>
// (the role of cpprob::tuple_observes_t, /root/reference include/cpprob/metapriors.hpp)
template<class F> struct observes_of;
template<class... A> struct observes_of<void (*)(A...)> { using type = std::tuple<std::decay_t<A>...>; };

template<class F>
void execute(const F & model, const options & opt)
{
    const std::string folder = opt.model_folder.empty() ? opt.model : opt.model_folder;
    ::mkdir(folder.c_str(), 0777);
    const std::string post_file_sis = folder + "/" + opt.generated_file + "_sis";
    if (opt.sis) {
        if (opt.observes_file.empty() == opt.observes.empty()) {
            std::cerr << R"(In CSIS or SIS mode exactly one of the options "--observes" or "--observes_file" has to be set)" << std::endl;
            std::exit(EXIT_FAILURE);
        }
        typename observes_of<F>::type observes;
        const bool ok = opt.observes_file.empty() ? cpprob::parse_string(opt.observes, observes)
                                                  : cpprob::parse_file(folder + "/" + opt.observes_file, observes);
        if (!ok) {
            std::cerr << "Could not parse the observations.\n"
                      << "Please use spaces to separate the observations and elements of an aggregate type instead of commas.\n"
                      << "If using the -o option, please surround the arguments by quotes.\n";
            std::exit(EXIT_FAILURE);
        }
        std::cout << "Sequential Importance Sampling (SIS)" << std::endl;
        cpprob::inference(cpprob::StateType::sis, model, observes, opt.n_samples, post_file_sis);
    }
    if (opt.estimate) {
        std::cout << "Posterior Distribution Estimators" << std::endl;
        std::cout << cpprob::StatsPrinter{post_file_sis};
    }
}

}  // namespace

int main(int argc, char ** argv)
{
    options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&]() -> std::string {
            if (i + 1 >= argc) {
                std::cerr << "option " << a << " needs a value\n";
                std::exit(EXIT_FAILURE);
            }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "--sis") opt.sis = true;
        else if (a == "--estimate") opt.estimate = true;
        else if (a == "--model") opt.model = value();
        else if (a == "--model_folder") opt.model_folder = value();
        else if (a == "-n" || a == "--n_samples") opt.n_samples = std::strtoull(value().c_str(), nullptr, 10);
        else if (a == "-o" || a == "--observes") opt.observes = value();
        else if (a == "-f" || a == "--observes_file") opt.observes_file = value();
        else if (a == "--generated_file") opt.generated_file = value();
        else if (a == "--compile" || a == "--csis" || a == "--dryrun") {
            std::cerr << a << ": inference compilation is not part of the B200 SIS engine; use the reference CPProb for it.\n";
            return EXIT_FAILURE;
        } else {
            std::cerr << "unknown option " << a << "\n";
            usage();
            return EXIT_FAILURE;
        }
    }
    if (opt.model.empty()) {
        std::cerr << "the option '--model' is required but missing\n";
        return EXIT_FAILURE;
    }
    try {
        // the registry of /root/reference src/main.cpp:123-130
        if (opt.model == "unk_mean") execute(&models::gaussian_unknown_mean<>, opt);
        else if (opt.model == "unk_mean_rejection") execute(&models::normal_rejection_sampling<>, opt);
        else if (opt.model == "linear_gaussian") execute(&models::linear_gaussian_1d<50>, opt);
        else if (opt.model == "hmm") execute(&models::hmm<10>, opt);
        else if (opt.model == "linear_regression") execute(&models::poly_adjustment<1, 6>, opt);
        else if (opt.model == "dyn_linear_reg") execute(&models::linear_regression<>, opt);
        else if (opt.model == "unk_mean_2d") execute(&models::gaussian_2d_unk_mean<>, opt);
        else {
            std::cerr << "Model not available. Please provide one of the following:" << std::endl << kModelNames << std::endl;
            return EXIT_FAILURE;
        }
    } catch (const std::exception & e) {
        std::cerr << "error: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
