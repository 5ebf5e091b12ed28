"""The oracle (CPU restatement of the reference SIS path) against every pin the reference offers:
its own known-answer grids for logpdf (tests/cpprob/logpdf.cpp), the survey's golden values, the
README / thesis posteriors — plus the analytic posteriors of the chain models."""
import math
import os
import re

import numpy as np
import pytest
from scipy import stats

import analytic

G = analytic.golden()
EPS_REF = 1e-8          # the reference's own tolerance, tests/cpprob/logpdf.cpp:16


def test_logpdf_normal_reference_grid(oracle):
    # full grid of tests/cpprob/logpdf.cpp:23-35 (both the log(pdf) and the exp(logpdf) assertion)
    for std in range(1, 20):
        rows = np.array([(mean / f, i) for mean in range(-10, 10) for f in range(1, 20) for i in range(-10, 10)])
        for mu in np.unique(rows[:, 0]):
            x = np.arange(-10, 10, dtype=float)
            got = oracle.logpdf("normal", [mu, std], x)
            np.testing.assert_allclose(got, stats.norm.logpdf(x, mu, std), atol=EPS_REF, rtol=0)
            np.testing.assert_allclose(np.exp(got), stats.norm.pdf(x, mu, std), atol=EPS_REF, rtol=0)


def test_logpdf_normal_golden_subsample(oracle):
    g = G["logpdf_normal"]
    for (mu, sd, x), exp in zip(g["mean_sigma_x"], g["expected"]):
        got = oracle.logpdf("normal", [mu, sd], [x])[0]
        assert abs(got - exp) <= 1e-12 * max(1.0, abs(exp))


def test_logpdf_uniform_reference_grid(oracle):
    # tests/cpprob/logpdf.cpp:61-78, including the -inf (out of support) cases
    g = G["logpdf_uniform_real"]
    for (a, b, x), exp in zip(g["a_b_x"], g["expected"]):
        got = oracle.logpdf("uniform_real", [a, b], [x])[0]
        if exp is None:
            assert got == -math.inf
        else:
            assert abs(got - exp) <= 1e-12 * max(1.0, abs(exp))


def test_survey_golden_values(oracle):
    for k in G["kat"]:
        got = oracle.logpdf(k["kind"], k["params"], [k["x"]])[0]
        assert abs(got - k["expected"]) <= 1e-14 * max(1.0, abs(k["expected"])), k


def test_logpdf_special_cases(oracle):
    inf = math.inf
    assert oracle.logpdf("normal", [1.0, 0.0], [1.0])[0] == 0.0            # sigma == 0: Dirac
    assert oracle.logpdf("normal", [1.0, 0.0], [1.5])[0] == -inf
    assert oracle.logpdf("normal", [0.0, 1.0], [inf])[0] == -inf           # |x| == inf
    assert oracle.logpdf("normal", [0.0, 1.0], [-inf])[0] == -inf
    assert oracle.logpdf("uniform_smallint", [0, 2], [3])[0] == -inf
    assert oracle.logpdf("discrete", [0.1, 0.5, 0.4], [3])[0] == -inf
    assert oracle.logpdf("discrete", [0.1, 0.5, 0.4], [-1])[0] == -inf
    assert oracle.logpdf("poisson", [0.0], [0])[0] == -inf                 # lambda == 0
    np.testing.assert_allclose(oracle.logpdf("poisson", [3.5], np.arange(0, 15.0)), stats.poisson.logpmf(np.arange(0, 15), 3.5), rtol=1e-13)
    np.testing.assert_allclose(oracle.logpdf("discrete", [1.0, 5.0, 4.0], [0, 1, 2]), np.log([0.1, 0.5, 0.4]), rtol=1e-14)


def test_philox_known_answers(oracle):
    for k in G["philox4x32_10"]:
        assert oracle.philox(k["ctr"], k["key"])[0].tolist() == k["out"]


def test_readme_hello_world(oracle, tmp_path):
    """C1: gaussian_unknown_mean x=(3,4), 10,000 particles, faithful flavour; README.md:118 posterior."""
    prefix = str(tmp_path / "posterior_sis")
    n = 10_000
    oracle.run("gaussian_unknown_mean", [3.0, 4.0], n, prefix, how="faithful", seed=7)
    # files: .real and .ids only (.int / .any were written per trace, then removed by finish_infer)
    assert sorted(os.listdir(tmp_path)) == ["posterior_sis.ids", "posterior_sis.real"]
    assert open(prefix + ".ids").read() == "Mean\n"
    lines = open(prefix + ".real").read().splitlines()
    assert len(lines) == n
    num = r"-?\d\.\d{15}e[+-]\d{2}"
    pat = re.compile(rf"^\(\[\(0 ({num})\)\] ({num})\)$")
    m = pat.match(lines[0])
    assert m, lines[0]
    # the record's log-weight is the log-likelihood of (3,4) at the recorded mu
    mu, lw = float(m.group(1)), float(m.group(2))
    assert abs(lw - (stats.norm.logpdf(3, mu, 2) + stats.norm.logpdf(4, mu, 2))) < 1e-12
    ids, ks, mean, var = oracle.stats_real(prefix)
    r = G["readme_model"]
    se = 1.2973 / math.sqrt(n)
    assert abs(mean[0] - r["post_mean"]) < 4 * se
    assert abs(var[0] - r["post_var"]) < 0.15
    text = oracle.stats_text(prefix)
    assert text.startswith(f"Estimators for {prefix}.real\nMean:\n  Mean: ")
    assert "  Variance: " in text


def test_file_append_semantics(oracle, tmp_path):
    prefix = str(tmp_path / "p")
    oracle.run("gaussian_unknown_mean", [3.0, 4.0], 50, prefix, how="faithful")
    oracle.run("gaussian_unknown_mean", [3.0, 4.0], 70, prefix, how="fast")
    assert len(open(prefix + ".real").read().splitlines()) == 120      # ios::app, never truncated
    assert open(prefix + ".ids").read() == "Mean\n"                    # .ids is rewritten


def test_models_hpp_variant_thesis_posterior(oracle, tmp_path):
    prefix = str(tmp_path / "p")
    n = 40_000
    oracle.run("gaussian_unknown_mean_mu", G["models_hpp_variant"]["thesis_obs"], n, prefix, seed=3)
    assert open(prefix + ".ids").read() == "Mu\n"
    _, _, mean, var = oracle.stats_real(prefix)
    assert abs(mean[0] - 7.25) < 0.1 and abs(var[0] - 5 / 6) < 0.1      # prior far from the data: low ESS


def test_linear_gaussian_against_kalman(oracle, tmp_path):
    prefix = str(tmp_path / "lg")
    obs = G["obs_linear_gaussian_32"][:6]
    n = 60_000
    oracle.run("linear_gaussian_1d", obs, n, prefix, seed=11)
    ids, ks, mean, var = oracle.stats_real(prefix)
    assert ids.tolist() == [0] * 6 and ks.tolist() == list(range(6))
    ms, vs, _ = analytic.kalman_smoother(obs)
    np.testing.assert_allclose(mean, ms, atol=0.08)
    np.testing.assert_allclose(var, vs, atol=0.08)
    text = oracle.stats_text(prefix)
    assert "State 0:\n  Mean: " in text and "State 5:\n" in text         # k printed because the id repeats


def test_hmm_against_forward_backward(oracle, tmp_path):
    prefix = str(tmp_path / "hmm")
    obs = G["obs_hmm_64"][:8]
    n = 40_000
    oracle.run("hmm", obs, n, prefix, seed=5)
    assert sorted(os.listdir(tmp_path)) == ["hmm.ids", "hmm.int"]
    first = open(prefix + ".int").readline()
    assert re.match(r"^\(\[(\(0 [012]\) ){7}\(0 [012]\)\] -?\d\.\d{15}e[+-]\d{2}\)$", first), first
    ids, ks, prob, mp, npts = oracle.stats_int(prefix, 0, 3)
    post, _ = analytic.hmm_forward_backward(obs)
    np.testing.assert_allclose(prob, post, atol=0.02)
    assert (npts == n).all()
    assert (mp == post.argmax(1)).all()
    text = oracle.stats_text(prefix)
    assert "State 0:\n  Distribution:\n    0: " in text and "  MAP: " in text and f"  Num points: {n}" in text


def test_replay_matches_emitted_logw(oracle, tmp_path):
    prefix = str(tmp_path / "r")
    obs = G["obs_linear_gaussian_32"][:5]
    oracle.run("linear_gaussian_1d", obs, 200, prefix, seed=2)
    _, values, logw = oracle.parse_records(prefix + ".real", "real", 5, 1000)
    again = oracle.replay_logw("linear_gaussian_1d", obs, values)
    np.testing.assert_allclose(again, logw, rtol=1e-13)


def test_oracle_files_equal_the_reference_loops_committed_outputs(oracle, tmp_path):
    """tests/golden/ref_sis_golden.json: posterior files the REFERENCE'S OWN SIS loop wrote (oracle/_ref/ref_sis, generator
    tests/golden/make_ref_sis_golden.py) for prescribed sampled values.  The restated oracle, given the same values, must
    write the same bytes — this form of the pin needs neither /root/reference nor oracle/_ref at test time."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_sis_golden.json")))
    assert len(fx["cases"]) >= 6
    for i, c in enumerate(fx["cases"]):
        values = np.array([[float.fromhex(v) for v in row] for row in c["values_hex"]])
        prefix = str(tmp_path / f"c{i}")
        oracle.replay_files(c["model"], c["obs"], values, prefix)
        for ext in (".real", ".int", ".any", ".ids"):
            got = open(prefix + ext).read() if os.path.exists(prefix + ext) else None
            assert got == c["files"].get(ext), (c["model"], ext)
