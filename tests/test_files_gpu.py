"""Posterior files written by the GPU path: byte format, .ids, append / removal semantics, and
readability by the reference's parser grammar (restated in the oracle)."""
import os
import re

import numpy as np
import pytest

import analytic
from cpprob_b200 import capi

pytestmark = pytest.mark.gpu
G = analytic.golden()
NUM = r"(?:-?\d\.\d{15}e[+-]\d{2}|-?inf|-?nan)"


def test_real_file_format_and_ids(engine, oracle, tmp_path):
    prefix = str(tmp_path / "posterior_sis")
    n = 10_000
    st = engine.infer_to_files("gaussian_unknown_mean", [3.0, 4.0], n, prefix)
    assert sorted(os.listdir(tmp_path)) == ["posterior_sis.ids", "posterior_sis.real", "posterior_sis.stats"]
    assert open(prefix + ".ids").read() == "Mean\n"
    lines = open(prefix + ".real").read().splitlines()
    assert len(lines) == n
    pat = re.compile(rf"^\(\[\(0 {NUM}\)\] {NUM}\)$")
    assert all(pat.match(l) for l in lines)
    # readable by the reference grammar; values round-trip to 16 significant digits
    _, values, logw = oracle.parse_records(prefix + ".real", "real", 1, n)
    out = engine.run("gaussian_unknown_mean", [3.0, 4.0], n, collect=True)
    np.testing.assert_allclose(values[:, 0], out["real_rows"][0], rtol=6e-16)
    np.testing.assert_allclose(logw, out["log_w"], rtol=6e-16)
    # the formatter is byte-identical to printf("%.15e") == ostream << scientific << setprecision(15)
    for v, l in zip(out["real_rows"][0][:200], lines[:200]):
        assert l.startswith("([(0 %.15e)] " % v)
    # StatsPrinter (restated) on the GPU-written files prints the reference's text
    text = oracle.stats_text(prefix)
    assert text.startswith(f"Estimators for {prefix}.real\nMean:\n  Mean: {st['real_mean'][0]:.6g}"[:60])


def test_int_file_format(engine, oracle, tmp_path):
    prefix = str(tmp_path / "h")
    n = 3000
    engine.infer_to_files("hmm", G["obs_hmm_64"][:7], n, prefix)
    assert sorted(os.listdir(tmp_path)) == ["h.ids", "h.int", "h.stats"]
    lines = open(prefix + ".int").read().splitlines()
    pat = re.compile(rf"^\(\[(\(0 [012]\) ){{6}}\(0 [012]\)\] {NUM}\)$")
    assert len(lines) == n and all(pat.match(l) for l in lines)
    assert open(prefix + ".ids").read() == "State\n"


def test_append_and_removal_semantics(engine, tmp_path):
    prefix = str(tmp_path / "p")
    engine.infer_to_files("gaussian_unknown_mean", [3.0, 4.0], 100, prefix)
    engine.infer_to_files("gaussian_unknown_mean", [3.0, 4.0], 150, prefix)
    assert len(open(prefix + ".real").read().splitlines()) == 250        # ios::app (state.cpp:264)
    # a run without real predicts removes a pre-existing .real (finish_infer, state.cpp:167-175)
    engine.infer_to_files("hmm", G["obs_hmm_64"][:3], 64, prefix)
    assert not os.path.exists(prefix + ".real") and os.path.exists(prefix + ".int")
    assert open(prefix + ".ids").read() == "State\n"


def test_multi_batch_emission_order(tmp_path):
    """Small batches force the double-buffered copy pipeline through many rounds."""
    from cpprob_b200 import Engine
    with Engine(seed=0x5EED, max_batch=capi.CHUNK) as e:
        n = 5 * capi.CHUNK + 99
        out = e.run("linear_gaussian_1d", G["obs_linear_gaussian_32"][:4], n, collect=True)
    with Engine(seed=0x5EED) as e2:
        ref = e2.run("linear_gaussian_1d", G["obs_linear_gaussian_32"][:4], n, collect=True)
    assert (out["real_rows"] == ref["real_rows"]).all() and (out["log_w"] == ref["log_w"]).all()
    assert (out["sums"] == ref["sums"]).all()
