"""The restated oracle against the REFERENCE'S OWN SIS LOOP (oracle/_ref/ref_sis: cpprob::inference of cpprob.hpp:173-203 with
src/cpprob/{state,trace,sample,utils,serialization,cpprob,socket}.cpp and the reference's models linked unmodified; only
third-party Boost / FlatBuffers / ZeroMQ names are stood in for).  Both run on the SAME prescribed sampled values, so
everything downstream of a draw is compared: log-pdfs, the statement-by-statement log-weight accumulation, the routing of
predicts to .real / .int, first-seen address ids, the posterior-file bytes and the .ids file (SURVEY.md section 8 rows (a)5,
(a)6, (a)7)."""
import os

import numpy as np
import pytest

import analytic
import ref_lib

G = analytic.golden()
pytestmark = pytest.mark.skipif(not (ref_lib.available() and os.path.exists(ref_lib.REF_SIS)), reason="oracle/_ref/ref_sis not built and /root/reference absent")

PTS = [1, 2.1, 2, 3.9, 3, 5.3, 4, 7.7, 5, 10.2, 6, 12.9]      # the points of poly_adjustment.hpp:33
CASES = [  # (model, obs, sampled values per trace, kind of the sampled values, traces)
    ("gaussian_unknown_mean", [3.0, 4.0], 1, "real", 300),
    ("gaussian_unknown_mean_mu", [3.0, 4.0], 1, "real", 300),
    ("linear_gaussian_1d", G["obs_linear_gaussian_32"][:5], 5, "real", 200),
    ("linear_gaussian_1d", G["obs_linear_gaussian_32"][:8], 8, "real", 200),
    ("linear_gaussian_1d", G["obs_linear_gaussian_32"], 32, "real", 150),
    ("hmm", G["obs_hmm_64"][:9], 9, "state", 200),
    ("hmm", G["obs_hmm_64"][:12], 12, "state", 200),
    ("hmm", G["obs_hmm_64"], 64, "state", 150),
    ("hmm", G["obs_hmm_1000"], 1000, "state", 12),
    ("gaussian_2d_unk_mean", [1.5, 2.5], 2, "real", 200),
    ("poly_adjustment_1", PTS, 2, "real", 200),          # main.cpp's "linear_regression" = poly_adjustment<1, 6>
    ("poly_adjustment_2", PTS, 3, "real", 200),
    ("poly_adjustment_3", PTS, 4, "real", 200),
    ("linear_regression", PTS, 2, "real", 200),          # poly_adjustment.hpp:57-82, the model with a Builder argument
]


def read(path):
    return open(path, "rb").read() if os.path.exists(path) else None


@pytest.mark.parametrize("model,obs,per,kind,n", CASES, ids=[f"{c[0]}-{c[2]}" for c in CASES])
def test_oracle_files_equal_the_references_own_loop(oracle, tmp_path, model, obs, per, kind, n):
    rng = np.random.default_rng(per * 7 + n)
    values = rng.integers(0, 3, (n, per)).astype(np.float64) if kind == "state" else rng.normal(0.5, 2.0, (n, per))
    a, b = str(tmp_path / "oracle"), str(tmp_path / "ref")
    oracle.replay_files(model, obs, values, a)
    ref_lib.ref_sis(model, obs, n, b, replay=values)
    for ext in (".real", ".int", ".any", ".ids"):
        assert read(a + ext) == read(b + ext), ext
    assert read(b + ".ids") is not None and (read(b + ".real") is not None) != (read(b + ".int") is not None)


def test_rejection_sampling_loop_equals_the_references(oracle, tmp_path):
    """models.hpp:82-112: the prior is simulated by rejection inside `cpprob::rejection_sampling`.  Replay lists with a known
    fate — k proposals far in the tails paired with the largest acceptance draw (rejected whatever the pdf's last bit is),
    then one pair with acceptance draw 0 (accepted) — through the reference's loop and the restatement: same files."""
    rng = np.random.default_rng(5)
    n, values = 120, []
    for _ in range(n):
        for _ in range(int(rng.integers(0, 4))):
            values += [float(rng.choice([-43.0, 45.0])), 0.178]          # pdf there ~ 1e-84, maxval = 0.1784...: rejected
        values += [float(rng.normal(2.0, 1.5)), 0.0]                     # 0 > pdf is false: accepted
    values = np.array(values)
    a, b = str(tmp_path / "oracle"), str(tmp_path / "ref")
    oracle.replay_files("normal_rejection_sampling", [3.0, 4.0], values, a, n_traces=n)      # (ragged traces: a flat list)
    flat = values.reshape(1, -1)
    ref_lib.ref_sis("normal_rejection_sampling", [3.0, 4.0], n, b, replay=flat)
    for ext in (".real", ".int", ".any", ".ids"):
        assert read(a + ext) == read(b + ext), ext
    assert read(b + ".real").count(b"\n") == n and read(b + ".ids") == b"Mu\n"


def test_all_distr_values_and_log_weights_equal_the_references_ids_do_not(oracle, tmp_path):
    """src/models/models.cpp:13-47, one statement of every distribution, one-argument predicts.  On the same sampled values
    the restatement and the reference's loop agree on every value and on every log-weight (bit for bit: normal,
    uniform_smallint, uniform_real, poisson and diagonal multivariate-normal log-pdfs inside one trace).  They do NOT agree
    on addresses — a recorded deviation: the reference's get_addr() (utils.cpp:71-128, out of scope) carries the call site's
    offset when built with -rdynamic as its CMake does, so it numbers the five statements 0..4 and routes the non-const
    NDArray of the last predict to <file>.any; the restatement, like the device model, names the function once."""
    import re
    rng = np.random.default_rng(9)
    n = 150
    values = np.column_stack([rng.normal(1, 2, n), rng.integers(2, 8, n), rng.uniform(2, 9.5, n), rng.poisson(0.8, n),
                              rng.normal(1, 1.4, n), rng.normal(2, 1, n), rng.normal(3, 2.2, n), rng.normal(4, 1.7, n)]).astype(np.float64)
    a, b = str(tmp_path / "oracle"), str(tmp_path / "ref")
    oracle.replay_files("all_distr", [0.0, 0.0], values, a)
    ref_lib.ref_sis("all_distr", [0.0, 0.0], n, b, replay=values)
    ids = [i for i in open(b + ".ids").read().split("\n") if i]
    assert len(ids) == 5 and all(re.match(r"^\[models::all_distr\(int, int\)\+0x[0-9a-f]+\]$", i) for i in ids)
    assert open(a + ".ids").read() == "[models::all_distr(int, int)]\n"
    lw = lambda path: [l.rsplit(b" ", 1)[1] for l in open(path, "rb").read().splitlines()]
    assert lw(a + ".real") == lw(b + ".real") == lw(b + ".int") == lw(b + ".any") and len(lw(b + ".any")) == n
    num = rb"-?\d\.\d{15}e[+-]\d\d"
    real_a = [re.findall(num, l.rsplit(b"]", 1)[0]) for l in open(a + ".real", "rb").read().splitlines()]
    real_b = [re.findall(num, l.rsplit(b"]", 1)[0]) + re.findall(num, m.rsplit(b"]", 1)[0])
              for l, m in zip(open(b + ".real", "rb").read().splitlines(), open(b + ".any", "rb").read().splitlines())]
    assert real_a == real_b                                            # normal, uniform_real, then the four vector components
    ints = lambda path: [re.findall(rb"\(\d+ (\d+)\)", l) for l in open(path, "rb").read().splitlines()]
    assert ints(a + ".int") == ints(b + ".int")


def test_reference_loop_reproduces_the_readme_posterior(ref_sis_stats=None):
    """The reference's own loop drawing by itself (standard-library normals through its get_rng()): README.md:118 says mean
    2.32353, variance 1.05882 — the same pin the CUDA path is held to."""
    import tempfile
    ref = ref_lib.load()
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "posterior_sis")
        ref_lib.ref_sis("gaussian_unknown_mean", [3.0, 4.0], 200_000, prefix)
        text = ref.stats_text(prefix)
    import re
    m = re.search(r"Mean:\n  Mean: (\S+)\n  Variance: (\S+)", text)
    assert m, text
    assert abs(float(m.group(1)) - 2.323529411764706) < 4 * 1.2973 / np.sqrt(200_000) * 1.5
    assert abs(float(m.group(2)) - 1.0588235294117647) < 0.03
