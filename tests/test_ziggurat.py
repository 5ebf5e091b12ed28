"""The ziggurat standard-normal sampler (include/cpprob/random/philox.hpp, tables from tools/gen_ziggurat.py).
The reference draws normals with boost::random::normal_distribution from a random_device-seeded mt19937
(src/models/gaussian.cpp:10-12, include/cpprob/utils.hpp:34-42), so parity is distributional only
(SURVEY.md §8c).  Here: the table is what the generator produces, its layers have equal areas, and the host twin of
the sampler (the same header compiled for the CPU) passes chi-square / tail / moment checks."""
import math
import os
import re
import subprocess

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE = os.path.join(ROOT, "include", "cpprob", "random", "ziggurat_table.inc")


def table():
    text = open(TABLE).read().replace("\\\n", " ")
    rows = {}
    for name in ("X", "F"):
        m = re.search(rf"#define CPPROB_ZIG_{name}_ROWS\s+(.*)", text)
        rows[name] = np.array([float(v) for v in m.group(1).split(",") if v.strip()])
    n = int(re.search(r"#define CPPROB_ZIG_N (\d+)", text).group(1))
    r = float(re.search(r"#define CPPROB_ZIG_R (\S+)", text).group(1))
    return n, r, rows["X"], rows["F"]


def test_table_is_reproducible(tmp_path):
    out = str(tmp_path / "zig.inc")
    subprocess.run(["python", os.path.join(ROOT, "tools", "gen_ziggurat.py"), "8192", out], check=True, capture_output=True)
    assert open(out).read() == open(TABLE).read()


def test_layers_have_equal_area():
    n, r, x, f = table()
    assert n == 8192 and len(x) == n + 1 and len(f) == n + 1
    assert x[1] == r and x[n] == 0.0 and f[n] == 1.0 and (np.diff(x) < 0).all()
    np.testing.assert_allclose(f[1:], np.exp(-0.5 * x[1:] ** 2), rtol=4e-15)   # x rounded to double moves f by x^2 ulp
    v = r * math.exp(-0.5 * r * r) + math.sqrt(math.pi / 2) * math.erfc(r / math.sqrt(2))
    np.testing.assert_allclose(x[1:n] * (f[2:] - f[1:n]), v, rtol=2e-11)     # rectangles 1..N-1
    np.testing.assert_allclose(x[0] * f[1], v, rtol=1e-14)                   # base strip incl. the tail
    # N layers x area V x 2 sides cover sqrt(2 pi) with the expected overhead = 1 / acceptance
    accept = math.sqrt(2 * math.pi) / (2 * n * v)
    assert 0.9995 < accept < 1.0


def host_normals(seed, first, n, per=1):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples"), "bin/zig_check"], check=True)
    out = subprocess.run([os.path.join(ROOT, "examples", "bin", "zig_check"), str(seed), str(first), str(n), str(per)],
                         capture_output=True, check=True).stdout
    return np.frombuffer(out, dtype=np.float64)


def test_host_twin_distribution():
    n = 6_000_000
    s = host_normals(0x5EED, 0, n)
    assert len(s) == n and np.isfinite(s).all()
    edges = stats.norm.ppf(np.linspace(0, 1, 257)[1:-1])
    assert stats.chisquare(np.bincount(np.searchsorted(edges, s), minlength=256)).pvalue > 1e-4
    a = np.abs(s)
    for lo, hi in ((4.548600609949139, np.inf), (3.0, 4.548600609949139), (0.0, 0.01)):
        p = 2 * (stats.norm.sf(lo) - stats.norm.sf(hi))
        assert abs(np.count_nonzero((a >= lo) & (a < hi)) - n * p) < 5 * math.sqrt(n * p) + 1
    assert abs(s.mean()) < 5 / math.sqrt(n) and abs((s ** 2).mean() - 1) < 5 * math.sqrt(2 / n)
    assert abs((s ** 4).mean() - 3) < 5 * math.sqrt(96 / n)
    # many draws from one stream (linear_gaussian_1d / gamma use it that way): same law, no serial correlation
    t = host_normals(7, 0, 20_000, per=100).reshape(20_000, 100)
    assert stats.kstest(t.ravel(), stats.norm.cdf).pvalue > 1e-4
    assert abs(np.mean(t[:, :-1] * t[:, 1:])) < 5 / math.sqrt(t.size)
