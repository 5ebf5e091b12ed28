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


# ---- device-side text stage: the GPU formats the lines, only text crosses PCIe --------
def _files(tmp, name):
    return {ext: open(os.path.join(tmp, name + ext), "rb").read() for ext in (".real", ".int", ".ids")
            if os.path.exists(os.path.join(tmp, name + ext))}


@pytest.mark.parametrize("model,obs_key,n_obs,n", [
    ("gaussian_unknown_mean", None, 2, 70_001),
    ("linear_gaussian_1d", "obs_linear_gaussian_32", 32, 40_000),
    ("hmm", "obs_hmm_64", 64, 33_000),
    ("gaussian_2d_unk_mean", None, 2, 10_000),      # vector predicts: `(id [a b])`
    ("all_distr", None, 2, 20_000),                 # real and int kinds in the same run
])
def test_gpu_text_equals_host_text(model, obs_key, n_obs, n, tmp_path, monkeypatch):
    """The lines formatted on the GPU are byte-identical to the host writer's (std::to_chars == printf %.15e ==
    ostream << scientific << setprecision(15)), for every record kind, across several batches."""
    from cpprob_b200 import Engine
    obs = G[obs_key][:n_obs] if obs_key else ([3.0, 4.0][:n_obs] if n_obs else [])
    with Engine(seed=0xABCD, max_batch=capi.CHUNK) as e:
        if model not in e.models():
            pytest.skip(f"{model} not built in")
        try:
            e.describe(model, obs)
        except capi.SisError:
            pytest.skip(f"{model} takes other observations")
        monkeypatch.delenv("CPPROB_SIS_TEXT", raising=False)
        e.infer_to_files(model, obs, n, str(tmp_path / "gpu"))
        ts = e.text_stage_stats()
        monkeypatch.setenv("CPPROB_SIS_TEXT", "host")
        e.infer_to_files(model, obs, n, str(tmp_path / "host"))
    a, b = _files(tmp_path, "gpu"), _files(tmp_path, "host")
    assert a.keys() == b.keys() and len(a) >= 2
    for ext in a:
        assert a[ext] == b[ext], f"{model}{ext} differs between the GPU and the host formatter"
    assert ts["bytes"] == sum(len(v) for k, v in a.items() if k != ".ids") and ts["fixups"] == 0
    assert ts["kernel_ms"] > 0 and ts["copy_ms"] > 0


def test_gpu_text_special_values(tmp_path, monkeypatch):
    """-inf log-weights (an observation outside the support) print as `-inf`, as the reference's ostream does."""
    from cpprob_b200 import Engine
    with Engine(seed=3) as e:
        # linear_gaussian_1d with an infinite observation: logpdf<normal> returns -inf (utils_normal_distribution.hpp:30-33)
        obs = [0.5, float("inf"), 1.0]
        monkeypatch.delenv("CPPROB_SIS_TEXT", raising=False)
        e.infer_to_files("linear_gaussian_1d", obs, 1000, str(tmp_path / "gpu"))
        monkeypatch.setenv("CPPROB_SIS_TEXT", "host")
        e.infer_to_files("linear_gaussian_1d", obs, 1000, str(tmp_path / "host"))
    a, b = _files(tmp_path, "gpu"), _files(tmp_path, "host")
    assert a == b
    assert all(l.endswith(b" -inf)") for l in a[".real"].splitlines())


def test_gpu_text_ambiguous_records_are_fixed_on_the_host(tmp_path, monkeypatch):
    """Test hook: every 977th record is reported as undecidable and damaged on the device; the host re-formats
    exactly those from the device rows, so the file is still byte-identical."""
    from cpprob_b200 import Engine
    monkeypatch.setenv("CPPROB_SIS_TEXT_FORCE_AMBIGUOUS", "977")
    monkeypatch.delenv("CPPROB_SIS_TEXT", raising=False)
    n = 3 * capi.CHUNK + 5
    with Engine(seed=11, max_batch=capi.CHUNK) as e:
        e.infer_to_files("linear_gaussian_1d", G["obs_linear_gaussian_32"][:5], n, str(tmp_path / "gpu"))
        ts = e.text_stage_stats()
        e.infer_to_files("hmm", G["obs_hmm_64"][:9], n, str(tmp_path / "gpui"))
        ti = e.text_stage_stats()
    monkeypatch.delenv("CPPROB_SIS_TEXT_FORCE_AMBIGUOUS")
    monkeypatch.setenv("CPPROB_SIS_TEXT", "host")
    with Engine(seed=11) as e:
        e.infer_to_files("linear_gaussian_1d", G["obs_linear_gaussian_32"][:5], n, str(tmp_path / "host"))
        e.infer_to_files("hmm", G["obs_hmm_64"][:9], n, str(tmp_path / "hosti"))
    expected = (n + 976) // 977
    assert ts["fixups"] == expected and ti["fixups"] == expected
    assert _files(tmp_path, "gpu") == _files(tmp_path, "host")
    assert _files(tmp_path, "gpui") == _files(tmp_path, "hosti")
    assert b"#" not in _files(tmp_path, "gpu")[".real"]


# ---- multi-GPU emission: every rank writes its own particle range of the same files -------------------------------
@pytest.mark.parametrize("model,obs_key,n_obs,n", [
    ("gaussian_unknown_mean", None, 2, 7 * capi.CHUNK + 99),
    ("linear_gaussian_1d", "obs_linear_gaussian_32", 32, 3 * capi.CHUNK + 5),
    ("hmm", "obs_hmm_64", 64, 4 * capi.CHUNK + 1),
    ("all_distr", None, 2, 2 * capi.CHUNK + 7),          # .real and .int in one run
    ("gaussian_unknown_mean", None, 2, 1000),             # fewer chunks than ranks: some ranks own nothing
])
def test_multi_engine_files_equal_single_engine_files(model, obs_key, n_obs, n, tmp_path):
    """cpprob_sis_infer_to_files_multi with 3 ranks (on distinct GPUs when the box has them, else three engines on one
    GPU): <prefix>.real / .int / .ids are byte for byte what one engine writes, the estimators are the same bits, and a
    second run appends after the first."""
    import torch
    from cpprob_b200 import Engine
    obs = G[obs_key][:n_obs] if obs_key else [3.0, 4.0]
    n_dev = torch.cuda.device_count()
    engines = [Engine(device=r % n_dev, seed=0xF11E, max_batch=capi.CHUNK) for r in range(3)]
    try:
        multi = capi.infer_to_files_multi(engines, model, obs, n, str(tmp_path / "multi"))
        capi.infer_to_files_multi(engines, model, obs, n, str(tmp_path / "multi"))       # appended (ios::app, state.cpp:264)
    finally:
        for e in engines:
            e.close()
    with Engine(seed=0xF11E, max_batch=capi.CHUNK) as e:
        single = e.infer_to_files(model, obs, n, str(tmp_path / "single"))
    a, b = _files(tmp_path, "multi"), _files(tmp_path, "single")
    assert a.keys() == b.keys() and len(a) >= 2
    for ext in a:
        assert a[ext] == (b[ext] if ext == ".ids" else b[ext] + b[ext]), f"{model}{ext}"
    assert (multi["sums"] == single["sums"]).all()
    assert np.array_equal(multi["real_mean"], single["real_mean"]) and np.array_equal(multi["int_prob"], single["int_prob"])
    assert os.path.exists(tmp_path / "multi.stats")


def test_estimator_only_run_removes_stale_record_files(engine, tmp_path):
    """CPPROB_SIS_EMIT_NONE leaves .ids + .stats and no record file — not even an older run's, which StatsPrinter would
    otherwise print against the new .ids (it reads the record files first, stats_printer.hpp:38-39)."""
    prefix = str(tmp_path / "p")
    engine.infer_to_files("gaussian_unknown_mean", [3.0, 4.0], 500, prefix)
    assert os.path.exists(prefix + ".real")
    engine.infer_to_files("hmm", G["obs_hmm_64"][:4], 500, prefix, emit=capi.EMIT_NONE)
    assert sorted(os.listdir(tmp_path)) == ["p.ids", "p.stats"]
    assert open(prefix + ".ids").read() == "State\n"
