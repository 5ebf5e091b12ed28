"""The restated oracle pinned to the REFERENCE'S OWN CODE (oracle/_ref: serialization.hpp, ndarray.hpp,
empirical_distribution.hpp and stats_printer.hpp of /root/reference, compiled unmodified): posterior-file grammar (SURVEY.md
section 8 row (a)7) and StatsPrinter arithmetic / console text (row (a)8)."""
import os

import numpy as np
import pytest

import analytic
import ref_lib

G = analytic.golden()
pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built and /root/reference absent")

MODELS = [("gaussian_unknown_mean", [3.0, 4.0], "real", 1), ("linear_gaussian_1d", G["obs_linear_gaussian_32"], "real", 32),
          ("hmm", G["obs_hmm_64"], "int", 64), ("hmm", G["obs_hmm_1000"], "int", 1000)]


@pytest.fixture(scope="module")
def ref():
    return ref_lib.load()


def test_ref_is_the_reference(ref):
    assert "stats_printer.hpp" in ref.describe() and "/root/reference" in ref.describe()


def test_writer_known_lines(ref):
    # the README record (SURVEY.md section 8c golden value: log_w(mu = 2.0) = -3.849171427529236)
    assert ref.write_real([0], [2.345678901234567], -3.849171427529236) == b"([(0 2.345678901234567e+00)] -3.849171427529236e+00)\n"
    assert ref.write_real([], [], -1.5) == b"([] -1.500000000000000e+00)\n"
    assert ref.write_real([0, 0], [1.0, -2.5e-7], float("-inf")) == b"([(0 1.000000000000000e+00) (0 -2.500000000000000e-07)] -inf)\n"
    assert ref.write_int([0, 0, 1], [2, 0, 17], -0.25) == b"([(0 2) (0 0) (1 17)] -2.500000000000000e-01)\n"
    assert ref.write_ndarray([0], [2], [1.5, -2.0], -1.0) == b"([(0 [1.500000000000000e+00 -2.000000000000000e+00])] -1.000000000000000e+00)\n"


@pytest.mark.parametrize("model,obs,kind,per", MODELS)
def test_oracle_files_are_the_references_bytes(oracle, ref, tmp_path, model, obs, kind, per):
    """Every line the restated writer produced == the reference's operator<< on the values its own operator>> parsed."""
    prefix = str(tmp_path / "o")
    n = 300 if per < 1000 else 40
    oracle.run(model, obs, n, prefix, how="faithful", seed=23)
    lines = open(f"{prefix}.{kind}", "rb").read().splitlines()
    assert len(lines) == n
    for line in lines:
        again = ref.reprint(line, kind)
        assert again == line + b"\n"
    # and the restated parser reads what the reference's parser reads
    ids_r, vals_r, lw_r = ref.parse_file(f"{prefix}.{kind}", kind, per)
    ids_o, vals_o, lw_o = oracle.parse_records(f"{prefix}.{kind}", kind, per, n)
    assert (ids_r == ids_o).all() and (vals_r == vals_o).all() and (lw_r == lw_o).all()


def test_writer_agrees_on_random_values(oracle, ref, tmp_path):
    """The restated writer against the reference's on values spanning the double range (incl. subnormals, -inf)."""
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-300, 300, 200), [0.0, -0.0, 5e-324, 1.7976931348623157e308,
                          0.1, 1 / 3, 2.5, 1e15, 9.999999999999999e22, 0.30000000000000004]])
    for v in vals:
        line = ref.write_real([3], [v], -v if np.isfinite(v) else 0.0)
        r = ref.parse(line.rstrip(b"\n"), "real")
        if abs(v) < 1.7e308:        # DBL_MAX rounds up at 16 digits: the reference's own parser rejects its own line
            assert r is not None and r[0].tolist() == [3] and abs(r[1][0] - v) <= 6e-16 * abs(v)
        assert line == (b"([(3 %s)] %s)\n" % (b"%.15e" % v, b"%.15e" % (-v if np.isfinite(v) else 0.0)))


@pytest.mark.parametrize("model,obs,kind,per", MODELS[:3])
def test_oracle_stats_printer_is_the_references(oracle, ref, tmp_path, model, obs, kind, per):
    prefix = str(tmp_path / "s")
    n = 4000
    oracle.run(model, obs[:12], n, prefix, how="fast", seed=31)
    assert oracle.stats_text(prefix) == ref.stats_text(prefix)
    per = min(per, 12)
    _, vals, lw = ref.parse_file(f"{prefix}.{kind}", kind, per)
    if kind == "real":
        _, _, mean, var = oracle.stats_real(prefix)
        for k in range(per):
            m, v = ref.empirical_real(vals[:, k], lw)
            assert m == mean[k] and v == var[k]          # same arithmetic in the same order: bit-equal
    else:
        _, _, prob, mp, npts = oracle.stats_int(prefix, 0, 3)
        for k in range(per):
            d, m, npt = ref.empirical_int(vals[:, k], lw)
            assert [d.get(b, 0.0) for b in range(3)] == prob[k].tolist() and m == mp[k] and npt == npts[k] == n


def test_empirical_distribution_edge_cases(ref):
    # all weights -inf: exp(-inf - (-inf)) = NaN everywhere, as SURVEY.md section 7 "Edge semantics" says
    m, v = ref.empirical_real([1.0, 2.0], [float("-inf")] * 2)
    assert np.isnan(m) and np.isnan(v)
    # one dominant weight
    m, v = ref.empirical_real([1.0, 5.0], [0.0, -800.0])
    assert m == 1.0 and v == 0.0
    d, mp, npts = ref.empirical_int([2, 2, 0, 1], [0.0, 0.0, 0.0, np.log(2.0)])
    # two values tie at 0.4: std::max_element returns the first maximum in key order (empirical_distribution.hpp:47-50)
    assert mp == 1 and npts == 4 and abs(d[2] - 0.4) < 1e-15 and abs(d[1] - 0.4) < 1e-15 and abs(d[0] - 0.2) < 1e-15


def test_reference_parser_quirks(ref):
    """Behaviour of the reference's own reader that a drop-in has to know about (and does not have to share):
    `istream >> double` rejects "-inf", so StatsPrinter exits on a record with a -inf log-weight; NDArray's operator>>
    (ndarray.hpp:313-325) rejects a vector value that is followed by `)`, so it cannot read back the `(id [a b])` records
    the reference writes for vector-valued predicts."""
    assert ref.parse(b"([(0 1.000000000000000e+00)] -inf)", "real") is None
    assert ref.parse_ndarray_ok(b"([(0 1.500000000000000e+00)] -1.000000000000000e+00)") == 1
    assert ref.parse_ndarray_ok(b"([(0 [1.500000000000000e+00 -2.000000000000000e+00])] -1.000000000000000e+00)") == -1


LOGPDF_CASES = [  # rows (a)4c-4e: no reference test pins them (its "poisson" test re-tests the normal, tests/cpprob/logpdf.cpp:43-53)
    ("uniform_smallint", [0, 2], np.arange(-3, 7.0)),
    ("uniform_smallint", [-4, 11], np.arange(-8, 16.0)),
    ("uniform_smallint", [5, 5], np.arange(3, 8.0)),
    ("discrete", [0.1, 0.5, 0.4], np.arange(-2, 6.0)),
    ("discrete", [1.0, 5.0, 4.0, 2.0, 8.0], np.arange(-1, 7.0)),
    ("discrete", [3.0], np.arange(-1, 3.0)),
    ("discrete", [0.0, 2.0, 0.0, 6.0], np.arange(0, 4.0)),          # zero-weight categories: log(0) = -inf
    ("poisson", [0.8], np.arange(0, 40.0)),
    ("poisson", [0.0], np.arange(0, 3.0)),                           # lambda == 0: -inf everywhere (utils_poisson.hpp:28-31)
    ("poisson", [37.5], np.arange(0, 160.0)),
    ("poisson", [1e-3], np.arange(0, 12.0)),
    ("poisson", [2.0], np.arange(-3, 3.0)),                          # negative counts: the loop is empty, x*log(l) - l comes back
]


@pytest.mark.parametrize("kind,params,xs", LOGPDF_CASES)
def test_oracle_logpdfs_are_the_references(oracle, ref, kind, params, xs):
    """The restated uniform_smallint / discrete / poisson log-pdfs against the reference's own logpdf<> structs
    (utils_uniform_smallint.hpp:17-27, utils_discrete.hpp:17-27, utils_poisson.hpp:17-36), compiled unmodified: same
    operations in the same order, so the same bits."""
    got = oracle.logpdf(kind, params, xs)
    exp = ref.logpdf(kind, params, xs)
    assert got.tobytes() == exp.tobytes(), (got, exp)


def test_oracle_normal_and_uniform_logpdfs_are_the_references(oracle, ref):
    """Rows (a)4a / 4b on the reference's own test grid (tests/cpprob/logpdf.cpp:23-35: mu and x in -10..10 by 0.5, sigma
    0.5..10 by 0.5; :61-78 for the uniform) plus the special cases of utils_normal_distribution.hpp:28-36: same bits."""
    grid = np.arange(-10, 10.25, 0.5)
    xs = np.concatenate([grid, [np.inf, -np.inf, 1e-300, 1e300, -0.0]])
    for mu in grid:
        for sd in np.arange(0.5, 10.25, 0.5):
            assert oracle.logpdf("normal", [mu, sd], xs).tobytes() == ref.logpdf("normal", [mu, sd], xs).tobytes()
    for mu, sd in [(1.0, 0.0), (0.7, 2 ** 0.5), (-3.0, 1e-8), (2.0, 1e12)]:
        x = np.concatenate([xs, [mu]])
        assert oracle.logpdf("normal", [mu, sd], x).tobytes() == ref.logpdf("normal", [mu, sd], x).tobytes()
    for a, b in [(2.0, 9.5), (-1.0, 1.0), (0.0, 1e-9), (-1e6, 1e6)]:
        x = np.concatenate([np.linspace(a - 1, b + 1, 101), [a, b, np.nextafter(a, -np.inf), np.nextafter(b, np.inf)]])
        assert oracle.logpdf("uniform_real", [a, b], x).tobytes() == ref.logpdf("uniform_real", [a, b], x).tobytes()


@pytest.mark.parametrize("model,obs_key,per", [("gaussian_unknown_mean", None, 1), ("gaussian_unknown_mean_mu", None, 1),
                                                ("linear_gaussian_1d", "obs_linear_gaussian_32", 32), ("hmm", "obs_hmm_64", 64),
                                                ("hmm", "obs_hmm_1000", 1000), ("gaussian_2d_unk_mean", "2d", 2)])
def test_oracle_log_weights_are_the_references_logpdfs_accumulated(oracle, ref, model, obs_key, per):
    """Row (a)5: the oracle's replayed log-weights equal, BIT FOR BIT, the sum the reference's own logpdf<> code gives when
    accumulated statement by statement in program order."""
    rng = np.random.default_rng(11)
    obs = [3.0, 4.0] if obs_key is None else ([1.5, 2.5] if obs_key == "2d" else G[obs_key])
    n = 64 if per < 1000 else 8
    values = rng.integers(0, 3, (n, per)).astype(np.float64) if model == "hmm" else rng.normal(0.5, 2.0, (n, per))
    got = oracle.replay_logw(model, obs, values)
    exp = np.array([ref_lib.ref_log_w(ref, model, obs, v) for v in values])
    assert got.tobytes() == exp.tobytes(), np.abs(got - exp).max()
