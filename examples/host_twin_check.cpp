// Host-side checks that need no GPU (run by tests/test_host_api.py): the HD headers compile as plain
// C++14, the host twin of Philox reproduces the Random123 known answers, the structure probe reports the
// reference's address ids / occurrence indices, the serialization grammar round-trips, and a model
// stub called outside inference behaves as a dry run.
#include <array>
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "cpprob/cpprob.hpp"
#include "cpprob/serialization.hpp"
#include "models/models.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

int main()
{
    using namespace cpprob;
    {   // Philox4x32-10 known answers (Random123 kat_vectors)
        std::uint32_t o[4];
        philox4x32::block(0, 0, 0, 0, philox_keys(0u, 0u), o[0], o[1], o[2], o[3]);
        CHECK(o[0] == 0x6627e8d5u && o[1] == 0xe169c58du && o[2] == 0xbc57ac4cu && o[3] == 0x9b00dbd8u);
        philox4x32::block(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, philox_keys(0xffffffffu, 0xffffffffu), o[0], o[1], o[2], o[3]);
        CHECK(o[0] == 0x408f276du && o[1] == 0x41c83b0eu && o[2] == 0xa20bc7c6u && o[3] == 0x6d5451fdu);
        philox4x32::block(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, philox_keys(0xa4093822u, 0x299f31d0u), o[0], o[1], o[2], o[3]);
        CHECK(o[0] == 0xd16cfe09u && o[1] == 0x94fdccebu && o[2] == 0x5001e420u && o[3] == 0x24126ea1u);
        const philox_keys k64(0x299f31d0a4093822ull);
        CHECK(k64.k[0] == 0xa4093822u && k64.k[1] == 0x299f31d0u && k64.k[2] == 0xa4093822u + 0x9E3779B9u);
        CHECK(stream_of_particle(0) == 0 && stream_of_particle(255) == 255 && stream_of_particle(256) == 0 && stream_of_particle(512) == 256);
        CHECK(turn_of_particle(255) == 0 && turn_of_particle(256) == 1 && turn_of_particle(511) == 1 && turn_of_particle(512) == 0);
    }
    {   // structure probe == TraceInfer::register_addr_predict + StatsPrinter's (id, k) keys
        const double obs2[2] = {3, 4};
        auto s = probe_model(models::gaussian_unknown_mean_model{}, obs2, 2);
        CHECK(s.ids.size() == 1 && s.ids[0] == "Mean" && s.n_real == 1 && s.n_int == 0 && s.n_samples == 1);
        s = probe_model(models::gaussian_unknown_mean_mu_model{}, obs2, 2);
        CHECK(s.ids[0] == "Mu");
        const double obs5[5] = {0.1, 0.2, 0.3, 0.4, 0.5};
        s = probe_model(models::linear_gaussian_1d_model{}, obs5, 5);
        CHECK(s.ids.size() == 1 && s.ids[0] == "State" && s.n_real == 5 && s.slots[3].k == 3 && s.slots[3].row == 3 && !s.slots[3].is_int);
        s = probe_model(models::hmm_model{}, obs5, 5);
        CHECK(s.ids[0] == "State" && s.n_int == 5 && s.n_real == 0 && s.n_samples == 5 && s.slots[4].is_int && s.slots[4].k == 4);
        s = probe_model(models::gaussian_2d_unk_mean_model{}, obs2, 2);
        CHECK(s.ids[0] == "Mu" && s.slots.size() == 1 && s.slots[0].width == 2 && s.n_real == 2);
        s = probe_model(models::all_distr_model{}, obs2, 2);
        CHECK(s.ids.size() == 1 && s.ids[0] == "[models::all_distr(int, int)]" && s.slots.size() == 5 && s.slots[4].width == 4 &&
              s.slots[4].row == 2 && s.n_real == 6 && s.n_int == 2 && s.slots[3].is_int && s.slots[3].k == 1);
        const double pts[4] = {1, 2.1, 2, 3.9};
        s = probe_model(models::linear_regression_model{}, pts, 4);
        CHECK(s.ids.size() == 2 && s.ids[0] == "a" && s.ids[1] == "b");
        s = probe_model(models::poly_adjustment_model<3>{}, pts, 4);
        CHECK(s.n_real == 4 && s.slots[3].k == 3);
    }
    {   // log-pdfs on the host twin, survey golden values
        CHECK(std::abs(logpdf<normal_distribution<>>()(normal_distribution<>(1, 2), 3.0) - -2.112085713764618) < 1e-14);
        CHECK(std::abs(logpdf<poisson_distribution<>>()(poisson_distribution<>(0.8), 3) - -3.2611901231706844) < 1e-14);
        CHECK(std::abs(logpdf<uniform_real_distribution<>>()(uniform_real_distribution<>(2, 9.5), 5.0) - -2.0149030205422647) < 1e-14);
        CHECK(std::abs(logpdf<uniform_smallint<>>()(uniform_smallint<>(0, 2), 1) - -1.0986122886681098) < 1e-14);
        const double w[3] = {0.1, 0.5, 0.4};
        CHECK(std::abs(logpdf<discrete_distribution<>>()(discrete_distribution<>(w, w + 3), 1) - -0.6931471805599453) < 1e-14);
        CHECK(logpdf<normal_distribution<>>()(normal_distribution<>(1, 0), 1.0) == 0.0);
    }
    {   // serialization grammar (serialization.hpp of the reference)
        using rec_t = std::pair<std::vector<std::pair<std::size_t, double>>, double>;
        rec_t rec{{{0, 1.5}, {0, -2.0}}, -3.25};
        CHECK(text::to_string(rec) == "([(0 1.5) (0 -2)] -3.25)");
        CHECK(text::to_string(rec_t{{}, 0.5}) == "([] 0.5)");
        rec_t back;
        std::istringstream is("([(0 1.885250430283626e+00) (1 -7.5e-01)] -3.938525470754040e+00)");
        CHECK(text::io<rec_t>::read(is, back) && back.first.size() == 2 && back.first[1].first == 1 && back.second < -3.9);
        std::tuple<double, double> two;
        CHECK(parse_string("3 4", two) && std::get<0>(two) == 3 && std::get<1>(two) == 4);
        std::tuple<std::array<double, 3>> arr;
        CHECK(parse_string("[1.5 2 3]", arr) && std::get<0>(arr)[2] == 3);
        CHECK(!parse_string("[1.5 2", arr));
        std::map<int, double> m{{1, 0.5}, {2, 0.25}};
        CHECK(text::to_string(m) == "{(1 0.5) (2 0.25)}");
        std::tuple<int, std::vector<int>> tv{7, {1, 2}};
        CHECK(text::to_string(tv) == "(7 [1 2])");
    }
    {   // outside inference a model stub is a dry run; inference on an unbound callable throws
        models::gaussian_unknown_mean<>(3.0, 4.0);
        models::normal_rejection_sampling<>(3.0, 4.0);
        models::gaussian_2d_unk_mean<>(std::vector<double>{3.0, 4.0});
        models::all_distr(0, 0);
        models::poly_adjustment<2, 2>(std::array<std::array<double, 2>, 2>{{{{1, 2.1}}, {{2, 3.9}}}});
        models::linear_regression<>(std::vector<std::pair<double, double>>{{1, 2.1}, {2, 3.9}});
        bool threw = false;
        try {
            cpprob::inference(cpprob::StateType::sis, [](double, double) {}, std::make_tuple(3., 4.), 10, "/tmp/never");
        } catch (const std::runtime_error & e) {
            threw = std::string(e.what()).find("no CPU fallback") != std::string::npos;
        }
        CHECK(threw);
        threw = false;
        try {
            cpprob::inference(cpprob::StateType::csis, &models::gaussian_unknown_mean<>, std::make_tuple(3., 4.), 10, "/tmp/never");
        } catch (const std::runtime_error &) { threw = true; }
        CHECK(threw);
        const double x = cpprob::sample(cpprob::normal_distribution<>(0, 1), true);
        CHECK(x == x);
    }
    std::printf(failures ? "host twin check: %d failure(s)\n" : "host twin check: ok\n", failures);
    return failures ? 1 : 0;
}
