"""A user model compiled as a plugin (examples/coin_plugin.cu): its own nvcc-built shared object
registers itself with the engine when it is loaded.  CPU: the plugin builds against the public headers and registers.
GPU: theta ~ Beta(2, 2) with Bernoulli(theta) flips gives the conjugate posterior Beta(2 + heads, 2 + tails)."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from cpprob_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "examples", "bin", "libcoin_plugin.so")


def load_plugin():
    env = dict(os.environ)
    if "/usr/local/cuda/bin" not in env.get("PATH", ""):
        env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples"), "bin/libcoin_plugin.so"], check=True, env=env)
    capi.lib()                                         # the engine first: the plugin links against it
    return ctypes.CDLL(PLUGIN, mode=ctypes.RTLD_GLOBAL)


def test_plugin_builds_and_registers():
    L = capi.lib()
    before = L.cpprob_sis_model_count()
    load_plugin()
    assert L.cpprob_sis_find_model(b"coin") >= 0
    assert L.cpprob_sis_model_count() >= before        # registering twice (another test loaded it) is idempotent by name


@pytest.mark.gpu
def test_plugin_posterior_is_conjugate(engine):
    load_plugin()
    flips = np.array([1, 1, 0, 1, 1, 1, 0, 1, 0, 1, 1, 1], dtype=np.float64)      # 9 heads, 3 tails
    d = engine.describe("coin", flips)
    assert d["ids"] == ["Theta"] and d["n_real"] == 1 and d["n_int"] == 0
    n = 4_000_000
    st = engine.run("coin", flips, n)
    a, b = 2 + 9, 2 + 3
    mean, var = a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1))
    se = math.sqrt(var / st["ess"])
    assert abs(st["real_mean"][0] - mean) < 5 * se
    assert abs(st["real_var"][0] - var) < 0.02 * var
    # log-evidence: B(a, b) / B(2, 2)
    log_z = (math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)) - (2 * math.lgamma(2) - math.lgamma(4))
    assert abs(st["log_evidence"] - log_z) < 5 * math.sqrt((n / st["ess"] - 1) / n)
    # the row path gives the same estimators
    rows = engine.run("coin", flips, n, force_rows=True)
    np.testing.assert_allclose(rows["real_mean"], st["real_mean"], rtol=1e-10)


@pytest.mark.gpu
def test_declared_int_range(engine):
    """Model::int_predict_states: a true declaration selects the packed staged path (no pilot for the window) and agrees with
    the row path bit for bit; a false one ends the run with CPPROB_SIS_ERANGE on every path instead of dropping values."""
    load_plugin()
    rolls = np.array([0.2, 2.9, 1.1, 3.3, 0.7, 2.2, 1.9])
    n = 3 * capi.CHUNK + 77
    a = engine.run("dice4", rolls, n)
    b = engine.run("dice4", rolls, n, force_rows=True)
    assert a["path"] == "staged" and b["path"] == "rows"
    assert a["int_lo"] == 0 and a["int_bins"] == 4 and (a["sums"] == b["sums"]).all()
    np.testing.assert_allclose(a["int_prob"].sum(1), 1.0, rtol=1e-12)
    # exact posterior of each face: independent rolls, p(face | r) ~ exp(-(r - face)^2 / 2)
    for k, r in enumerate(rolls):
        p = np.exp(-0.5 * (r - np.arange(4)) ** 2)
        np.testing.assert_allclose(a["int_prob"][k], p / p.sum(), atol=5.0 / math.sqrt(a["ess"]))
    for kw in ({}, {"force_rows": True}):
        with pytest.raises(capi.SisError) as ei:
            engine.run("dice_lying", rolls, n, **kw)
        assert ei.value.code == -5 and "declares" in str(ei.value)
