"""The CUDA path against the REFERENCE'S OWN CODE (oracle/_ref = serialization.hpp, ndarray.hpp, empirical_distribution.hpp,
stats_printer.hpp of the reference, compiled unmodified; it travels to the GPU box as a built file):
  * the posterior files the GPU writes parse with the reference's operator>> and its operator<< reproduces every line byte
    for byte, for all five BASELINE configs;
  * the on-device estimators equal EmpiricalDistribution on identical records (row (a)8);
  * the reference's StatsPrinter prints the device estimators from the GPU-written files."""
import re

import numpy as np
import pytest

import analytic
import ref_lib
from cpprob_b200 import capi

G = analytic.golden()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref was not built")]

CONFIGS = [  # (label, model, obs, kind, predicts per record, particles)
    ("C1", "gaussian_unknown_mean", [3.0, 4.0], "real", 1, 10_000),
    ("C2", "gaussian_unknown_mean", [3.0, 4.0], "real", 1, 3 * capi.CHUNK + 17),
    ("C3", "linear_gaussian_1d", G["obs_linear_gaussian_32"], "real", 32, 2 * capi.CHUNK + 5),
    ("C4", "hmm", G["obs_hmm_64"], "int", 64, 2 * capi.CHUNK + 5),
    ("C5", "hmm", G["obs_hmm_1000"], "int", 1000, 4100),
]


@pytest.fixture(scope="module")
def ref():
    return ref_lib.load()


@pytest.mark.parametrize("label,model,obs,kind,per,n", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_gpu_files_are_the_references_grammar_byte_for_byte(engine, ref, tmp_path, label, model, obs, kind, per, n):
    prefix = str(tmp_path / label)
    engine.infer_to_files(model, obs, n, prefix)
    lines = open(f"{prefix}.{kind}", "rb").read().splitlines()
    assert len(lines) == n
    step = max(1, n // 400)                                    # ~400 lines spread over every batch, plus both ends
    for line in lines[:50] + lines[::step] + lines[-50:]:
        assert ref.reprint(line, kind) == line + b"\n"
    # the values the reference's parser reads are the values the engine generated (text carries 16 significant digits)
    out = engine.run(model, obs, n, collect=True)
    rows = out["real_rows"] if kind == "real" else out["int_rows"]
    pick = np.unique(np.concatenate([np.arange(0, n, step), [n - 1]]))
    for i in pick:
        ids, vals, lw = ref.parse(lines[i], kind, cap=per + 1)
        assert ids.size == per and (ids == 0).all()
        if kind == "real":
            np.testing.assert_allclose(vals, rows[:, i], rtol=6e-16, atol=0)
        else:
            assert (vals == rows[:, i]).all()
        assert abs(lw - out["log_w"][i]) <= 6e-16 * abs(out["log_w"][i])


@pytest.mark.parametrize("label,model,obs,kind,per,n", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_device_estimators_equal_the_references_empirical_distribution(engine, ref, label, model, obs, kind, per, n):
    """k_sis_fused / k_row_moments / k_row_hist / k_merge_columns against EmpiricalDistribution::mean, variance,
    distribution, max_a_posteriori on the very same (value, log_w) points."""
    out = engine.run(model, obs, n, collect=True)
    plain = engine.run(model, obs, n)                          # the estimator-only path (fused kernels) of the same run
    lw = out["log_w"]
    ks = range(per) if per <= 64 else list(range(0, per, 97)) + [per - 1]
    for st in (out, plain):
        for k in ks:
            if kind == "real":
                m, v = ref.empirical_real(out["real_rows"][k], lw)
                # same estimator, different summation order: agreement to rounding of sums of n terms
                assert abs(st["real_mean"][k] - m) <= 1e-11 * max(1.0, abs(m))
                assert abs(st["real_var"][k] - v) <= 1e-10 * max(1.0, abs(v))
            else:
                d, mp, npts = ref.empirical_int(out["int_rows"][k], lw)
                got = st["int_prob"][k]
                for b in range(int(st["int_bins"])):
                    assert abs(got[b] - d.get(int(st["int_lo"]) + b, 0.0)) <= 1e-11
                assert st["int_map"][k] == mp and npts == n
    # log-sum-exp as EmpiricalDistribution::logsumexp forms it (max-shifted)
    mx = lw.max()
    assert abs(plain["log_sum_exp"] - (mx + np.log(np.exp(lw - mx).sum()))) <= 1e-11 * abs(plain["log_sum_exp"])


@pytest.mark.parametrize("label,model,obs,kind,per,n", CONFIGS[:4], ids=[c[0] for c in CONFIGS[:4]])
def test_reference_stats_printer_prints_the_device_estimators(engine, ref, oracle, tmp_path, label, model, obs, kind, per, n):
    prefix = str(tmp_path / label)
    st = engine.infer_to_files(model, obs, n, prefix)
    text = ref.stats_text(prefix)                              # the reference's StatsPrinter on GPU-written files
    assert text == oracle.stats_text(prefix)                   # ... and the restated one prints the same
    num = r"(-?[0-9.]+(?:e[+-]\d+)?|-?nan|-?inf)"
    if kind == "real":
        got = re.findall(rf"  Mean: {num}\n  Variance: {num}\n", text)
        assert len(got) == per
        for k, (m, v) in enumerate(got):
            assert float(m) == pytest.approx(st["real_mean"][k], rel=2e-5, abs=1e-6)
            assert float(v) == pytest.approx(st["real_var"][k], rel=2e-5, abs=1e-6)
    else:
        maps = re.findall(r"  MAP: (\d+)\n  Num points: (\d+)\n", text)
        assert len(maps) == per
        assert [int(m) for m, _ in maps] == st["int_map"].tolist() and all(int(c) == n for _, c in maps)


@pytest.mark.parametrize("kind,params,xs", [
    ("uniform_smallint", [0, 2], np.arange(-3, 7.0)),
    ("uniform_smallint", [-4, 11], np.arange(-8, 16.0)),
    ("discrete", [0.1, 0.5, 0.4], np.arange(-2, 6.0)),
    ("discrete", [1.0, 5.0, 4.0, 2.0, 8.0], np.arange(-1, 7.0)),
    ("discrete", [0.0, 2.0, 0.0, 6.0], np.arange(0, 4.0)),
    ("poisson", [0.8], np.arange(0, 40.0)),
    ("poisson", [0.0], np.arange(0, 3.0)),
    ("poisson", [37.5], np.arange(0, 160.0)),
    ("poisson", [1e-3], np.arange(0, 12.0)),
])
def test_device_logpdfs_equal_the_references(engine, ref, kind, params, xs):
    """Rows (a)4c-4e: the device log-pdfs against the reference's own logpdf<> (utils_uniform_smallint.hpp:17-27,
    utils_discrete.hpp:17-27, utils_poisson.hpp:17-36 compiled unmodified).  Tolerance 1e-12 relative: the device `log` is
    this repo's own (<= 2 ulp), the sequence of operations is the reference's."""
    got = engine.logpdf(kind, params, xs)
    exp = ref.logpdf(kind, params, xs)
    assert (np.isneginf(got) == np.isneginf(exp)).all()
    fin = np.isfinite(exp)
    np.testing.assert_allclose(got[fin], exp[fin], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("model,obs_key,per", [("gaussian_unknown_mean", None, 1), ("linear_gaussian_1d", "obs_linear_gaussian_32", 32),
                                                ("hmm", "obs_hmm_64", 64), ("hmm", "obs_hmm_1000", 1000),
                                                ("gaussian_2d_unk_mean", "2d", 2)])
def test_device_log_weights_equal_the_references_logpdfs_accumulated(engine, ref, model, obs_key, per):
    """Row (a)5 against the reference's own code: the log-weight the device computes for a given trace equals the sum of the
    reference's logpdf<> values (utils_normal_distribution.hpp:20-45, utils_multivariate_normal.hpp:20-33 compiled
    unmodified) accumulated statement by statement, within 1e-12 relative (the device `log` is this repo's own, <= 2 ulp)."""
    rng = np.random.default_rng(12)
    obs = [3.0, 4.0] if obs_key is None else ([1.5, 2.5] if obs_key == "2d" else G[obs_key])
    n = 256 if per < 1000 else 16
    if model == "hmm":
        values = rng.integers(0, 3, (n, per))
        got = engine.replay(model, obs, int_rows=values.T.astype(np.int32))
    else:
        values = rng.normal(0.5, 2.0, (n, per))
        got = engine.replay(model, obs, real_rows=values.T)
    exp = np.array([ref_lib.ref_log_w(ref, model, obs, v) for v in values])
    np.testing.assert_allclose(got, exp, rtol=1e-12)
