"""Device distribution layer on the GPU (through the C ABI): Philox known answers, fp64 elementary
functions, log-pdfs against the oracle / the reference's grids, samplers against scipy."""
import math

import mpmath
import numpy as np
import pytest
from scipy import stats

import analytic

pytestmark = pytest.mark.gpu
G = analytic.golden()
mpmath.mp.dps = 40


def ulp_err(got, exact):
    exact = np.asarray(exact, dtype=np.float64)
    return np.abs(got - exact) / np.spacing(np.abs(exact))


def test_philox_known_answers(engine, oracle):
    for k in G["philox4x32_10"]:
        assert engine.philox(k["ctr"], k["key"])[0].tolist() == k["out"]
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, size=(4096, 2), dtype=np.uint64).astype(np.uint32)
    assert (engine.philox(ctr, key) == oracle.philox(ctr, key)).all()          # bit-exact


def test_log_unit(engine):
    rng = np.random.default_rng(1)
    u = np.concatenate([rng.random(200_000), 2.0 ** -rng.uniform(0, 52, 50_000), [1 - 2 ** -53, 2 ** -53, 0.5, 0.70710678, 0.7071068]])
    got = engine.dmath(0, u)
    exact = np.array([float(mpmath.log(mpmath.mpf(x))) for x in u[:20_000]])
    assert ulp_err(got[:20_000], exact).max() <= 2.0
    assert ulp_err(got, np.log(u)).max() <= 3.0


def test_log_general(engine):
    x = np.array([0.0, -1.0, math.inf, math.nan, 5e-324, 1e-310, 2.2250738585072014e-308, 1.0, 2 * math.pi * 4, 7.5, 1e300])
    got = engine.dmath(6, x)
    assert got[0] == -math.inf and math.isnan(got[1]) and got[2] == math.inf and math.isnan(got[3])
    assert ulp_err(got[4:], np.log(x[4:])).max() <= 2.0
    r = np.random.default_rng(2).uniform(-700, 700, 100_000)
    assert ulp_err(engine.dmath(6, np.exp(r)), np.log(np.exp(r))).max() <= 2.0


def test_exp_weight(engine):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-707.9, 707.9, 200_000), rng.uniform(-40, 5, 200_000), [0.0, -0.0, 1e-300, -1e-17]])
    got = engine.dmath(1, x)
    exact = np.array([float(mpmath.exp(mpmath.mpf(v))) for v in x[:20_000]])
    assert ulp_err(got[:20_000], exact).max() <= 2.0
    assert ulp_err(got, np.exp(x)).max() <= 3.0
    # -inf and everything below the normal range flush to exactly 0 (weights of impossible traces);
    # +overflow / NaN are outside exp_weight's contract (the engine re-bases / poisons instead)
    special = engine.dmath(1, [-math.inf, -1e4, -745.2, -709.5, -708.0])
    assert special[:4].tolist() == [0.0, 0.0, 0.0, 0.0]
    assert abs(special[4] / math.exp(-708.0) - 1) < 1e-15


def test_exp_weight_tab(engine):
    """The fused kernel's table-assisted exp (2^(j/4096) table + degree-3 tail, 8 FP64 instructions): <= 2 ulp wherever
    its contract holds (finite argument, normal result); NaN for non-finite arguments (which is what poisons a unit's
    weight sum and sends it to the careful pass)."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-707.0, 707.0, 200_000), rng.uniform(-40, 5, 200_000), rng.uniform(-1e-3, 1e-3, 1000),
                        [0.0, -0.0, 1e-300, -1e-17, math.log(2) / 8192, -math.log(2) / 8192, math.log(2) / 4096]])
    got = engine.dmath(9, x)
    exact = np.array([float(mpmath.exp(mpmath.mpf(v))) for v in x[-25_000:]])
    assert ulp_err(got[-25_000:], exact).max() <= 2.0
    assert ulp_err(got, np.exp(x)).max() <= 3.0
    assert np.isnan(engine.dmath(9, [-math.inf, math.inf, math.nan])).all()


def test_sincos_2pi(engine):
    rng = np.random.default_rng(4)
    u = np.concatenate([rng.random(200_000), [0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0, 2 ** -53, 1 - 2 ** -53]])
    mp_sin = np.array([float(mpmath.sin(2 * mpmath.pi * mpmath.mpf(v))) for v in u[-4000:]])
    mp_cos = np.array([float(mpmath.cos(2 * mpmath.pi * mpmath.mpf(v))) for v in u[-4000:]])
    for fn, exact in ((2, mp_sin), (8, mp_sin), (3, mp_cos), (7, mp_cos)):
        got = engine.dmath(fn, u)
        assert np.abs(got[-4000:] - exact).max() <= 2.0 ** -52
        assert np.abs(got).max() <= 1.0
    # the joint and the single-chain evaluations agree to the last bit or two
    assert np.abs(engine.dmath(2, u) - engine.dmath(8, u)).max() <= 2.0 ** -52
    assert np.abs(engine.dmath(3, u) - engine.dmath(7, u)).max() <= 2.0 ** -52


def test_sqrt_pos(engine):
    x = np.concatenate([np.random.default_rng(5).uniform(1e-16, 80, 200_000), [2.2e-16, 1.0, 2.0, 4.0, 72.1]])
    assert ulp_err(engine.dmath(4, x), np.sqrt(x)).max() <= 1.0


@pytest.mark.parametrize("kind,params,xs", [
    ("normal", [1.0, 2.0], np.linspace(-30, 30, 2001)),
    ("normal", [0.7, 2 ** 0.5], np.array([-2.3, 0.0, math.inf, -math.inf])),
    ("normal", [1.0, 0.0], np.array([1.0, 1.5])),
    ("uniform_real", [2.0, 9.5], np.linspace(0, 12, 121)),
    ("uniform_smallint", [0, 2], np.arange(-2, 6.0)),
    ("discrete", [0.1, 0.5, 0.4], np.arange(-1, 5.0)),
    ("discrete", [1.0, 5.0, 4.0, 2.0, 8.0], np.arange(-1, 7.0)),
    ("poisson", [0.8], np.arange(0, 30.0)),
    ("poisson", [0.0], np.arange(0, 3.0)),
    ("poisson", [37.5], np.arange(0, 120.0)),
])
def test_logpdf_matches_oracle(engine, oracle, kind, params, xs):
    got = engine.logpdf(kind, params, xs)
    exp = oracle.logpdf(kind, params, xs)
    assert (np.isneginf(got) == np.isneginf(exp)).all()
    fin = np.isfinite(exp)
    np.testing.assert_allclose(got[fin], exp[fin], rtol=1e-12, atol=1e-300)


def test_logpdf_reference_grids(engine):
    g = G["logpdf_normal"]
    for (mu, sd, x), exp in list(zip(g["mean_sigma_x"], g["expected"]))[::10]:
        got = engine.logpdf("normal", [mu, sd], [x])[0]
        assert abs(got - exp) <= 1e-12 * max(1.0, abs(exp))
    # whole sigma-slices of the reference grid (tests/cpprob/logpdf.cpp:23-35), eps = 1e-8 there
    x = np.arange(-10, 10, dtype=float)
    for sd in (1, 7, 19):
        for mu in (-10.0, -10 / 19, 0.0, 9 / 7, 9.0):
            got = engine.logpdf("normal", [mu, sd], x)
            np.testing.assert_allclose(got, stats.norm.logpdf(x, mu, sd), rtol=1e-12)
            np.testing.assert_allclose(np.exp(got), stats.norm.pdf(x, mu, sd), atol=1e-8)
    g = G["logpdf_uniform_real"]
    for (a, b, x), exp in list(zip(g["a_b_x"], g["expected"]))[::10]:
        got = engine.logpdf("uniform_real", [a, b], [x])[0]
        assert (got == -math.inf) if exp is None else abs(got - exp) <= 1e-12 * max(1.0, abs(exp))
    for k in G["kat"]:
        got = engine.logpdf(k["kind"], k["params"], [k["x"]])[0]
        assert abs(got - k["expected"]) <= 1e-12 * abs(k["expected"])


def test_logpdf_gamma_beta_vs_scipy(engine):
    x = np.linspace(0.01, 20, 500)
    np.testing.assert_allclose(engine.logpdf("gamma", [2.5, 1.7], x), stats.gamma.logpdf(x, 2.5, scale=1.7), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(engine.logpdf("gamma", [0.6, 3.0], x), stats.gamma.logpdf(x, 0.6, scale=3.0), rtol=1e-11, atol=1e-12)
    assert engine.logpdf("gamma", [2.0, 1.0], [-1.0])[0] == -math.inf
    x = np.linspace(0.001, 0.999, 500)
    np.testing.assert_allclose(engine.logpdf("beta", [2.0, 5.0], x), stats.beta.logpdf(x, 2, 5), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(engine.logpdf("beta", [0.5, 0.5], x), stats.beta.logpdf(x, 0.5, 0.5), rtol=1e-11, atol=1e-12)
    assert engine.logpdf("beta", [2.0, 5.0], [1.5])[0] == -math.inf


N_S = 400_000


def ks_ok(sample, cdf):
    return stats.kstest(sample, cdf).pvalue > 1e-4


def test_sampler_normal(engine):
    s = engine.sample("normal", [1.0, 1.5], N_S, seed=11)
    assert ks_ok(s, stats.norm(1.0, 1.5).cdf)
    assert abs(s.mean() - 1.0) < 5 * 1.5 / math.sqrt(N_S)
    # reproducible and keyed by (seed, stream): a shifted window overlaps exactly
    s2 = engine.sample("normal", [1.0, 1.5], 1000, seed=11, first=500)
    assert (s2[:500] == s[500:1000]).all()
    assert not (engine.sample("normal", [1.0, 1.5], 1000, seed=12) == s[:1000]).any()


def test_sampler_normal_ziggurat_is_exact(engine):
    """The ziggurat (include/cpprob/random/philox.hpp) at 6.4e7 draws: equiprobable-bin chi-square, the tail beyond
    r = 4.5486 (only produced by the slow path's exponential rejection), the region around the layer edges, the
    moments up to the 4th and the sign balance, all within 5 sigma of N(0,1)."""
    n = 64_000_000
    s = engine.sample("normal", [0.0, 1.0], n, seed=20240607)
    assert np.isfinite(s).all()
    edges = stats.norm.ppf(np.linspace(0, 1, 513)[1:-1])
    counts = np.bincount(np.searchsorted(edges, s), minlength=512)
    assert stats.chisquare(counts).pvalue > 1e-4
    a = np.abs(s)
    for lo, hi in ((4.548600609949139, np.inf), (3.0, 4.548600609949139), (4.8, np.inf), (0.0, 0.01), (4.4, 4.6)):
        p = 2 * (stats.norm.sf(lo) - stats.norm.sf(hi))
        got = np.count_nonzero((a >= lo) & (a < hi))
        assert abs(got - n * p) < 5 * math.sqrt(n * p) + 1, (lo, hi, got, n * p)
    assert abs(s.mean()) < 5 / math.sqrt(n)
    assert abs((s ** 2).mean() - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs((s ** 3).mean()) < 5 * math.sqrt(15.0 / n)
    assert abs((s ** 4).mean() - 3.0) < 5 * math.sqrt(96.0 / n)
    assert abs(np.count_nonzero(s > 0) - n / 2) < 5 * math.sqrt(n / 4)
    # finer than any layer: a chi-square inside the innermost 1 % (the top layers, all wedge / cap draws)
    inner = s[a < stats.norm.ppf(0.505)]
    c2 = np.histogram(inner, bins=64)[0]
    e2 = np.diff(stats.norm.cdf(np.linspace(-stats.norm.ppf(0.505), stats.norm.ppf(0.505), 65)))
    assert stats.chisquare(c2, e2 / e2.sum() * c2.sum()).pvalue > 1e-4


def test_sampler_normal_matches_host_twin(engine):
    """Same header, same bits: the GPU's normals equal the host twin's (examples/zig_check.cpp).  The fast path is
    integer work plus one fma; only the 0.06 % slow-path draws call exp / log, where libm and the device may differ
    in the last place."""
    import test_ziggurat
    n = 3_000_000                                          # ~16 draws beyond r expected
    host = test_ziggurat.host_normals(99, 1000, n)
    dev = engine.sample("normal", [0.0, 1.0], n, seed=99, first=1000)
    same = host == dev
    assert same.mean() > 0.9999
    np.testing.assert_allclose(dev[~same], host[~same], rtol=1e-14)
    assert np.abs(dev).max() > 4.548600609949139          # the tail branch was exercised


def test_sampler_uniform_real(engine):
    s = engine.sample("uniform_real", [2.0, 9.5], N_S, seed=3)
    assert s.min() > 2.0 and s.max() < 9.5 and ks_ok(s, stats.uniform(2.0, 7.5).cdf)


def chi2_ok(counts, probs):
    n = counts.sum()
    return stats.chisquare(counts, probs * n).pvalue > 1e-4


def test_sampler_discrete_family(engine):
    s = engine.sample("uniform_smallint", [2, 7], N_S, seed=5).astype(int)
    assert s.min() == 2 and s.max() == 7 and chi2_ok(np.bincount(s - 2, minlength=6), np.full(6, 1 / 6))
    w = np.array([1.0, 5.0, 4.0, 0.5])
    s = engine.sample("discrete", w, N_S, seed=6).astype(int)
    assert chi2_ok(np.bincount(s, minlength=4), w / w.sum())


@pytest.mark.parametrize("lam", [0.8, 4.2, 12.0, 150.0])
def test_sampler_poisson(engine, lam):
    s = engine.sample("poisson", [lam], N_S, seed=7).astype(int)
    assert abs(s.mean() - lam) < 5 * math.sqrt(lam / N_S) and abs(s.var() - lam) < 0.05 * lam + 0.02
    top = int(lam + 6 * math.sqrt(lam) + 6)
    counts = np.bincount(np.minimum(s, top), minlength=top + 1)
    probs = stats.poisson.pmf(np.arange(top + 1), lam)
    probs[-1] += stats.poisson.sf(top, lam)
    keep = probs * N_S > 5
    merged = np.append(counts[keep], counts[~keep].sum())
    mp = np.append(probs[keep], probs[~keep].sum())
    if mp[-1] * N_S < 1:
        merged, mp = merged[:-1], mp[:-1] / mp[:-1].sum()
    assert chi2_ok(merged, mp / mp.sum())


@pytest.mark.parametrize("a,b", [(2.5, 1.7), (0.6, 3.0), (1.0, 1.0)])
def test_sampler_gamma(engine, a, b):
    s = engine.sample("gamma", [a, b], N_S, seed=8)
    assert s.min() > 0 and ks_ok(s, stats.gamma(a, scale=b).cdf)


@pytest.mark.parametrize("a,b", [(2.0, 5.0), (0.5, 0.5)])
def test_sampler_beta(engine, a, b):
    s = engine.sample("beta", [a, b], N_S, seed=9)
    assert 0 < s.min() and s.max() < 1 and ks_ok(s, stats.beta(a, b).cdf)
