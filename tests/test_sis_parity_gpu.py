"""Parity of the CUDA SIS path (through the C ABI) with the oracle restatement of the reference and
with the analytic posteriors.

Gates (BASELINE.json north_star):
  * replay: log-weights recomputed on the GPU from traces emitted by the (restated) reference match
    them to 1e-12 relative in fp64;
  * posterior moments agree with the reference's CPU SIS and with the analytic conjugate posterior
    (mean 2.32353, variance 1.05882 for x = (3, 4)) within 4 Monte-Carlo standard errors.
"""
import math

import numpy as np
import pytest

import analytic
from cpprob_b200 import capi

pytestmark = pytest.mark.gpu
G = analytic.golden()
RM = G["readme_model"]
OBS = {"gaussian_unknown_mean": [3.0, 4.0], "gaussian_unknown_mean_mu": [3.0, 4.0],
       "linear_gaussian_1d": G["obs_linear_gaussian_32"], "hmm": G["obs_hmm_64"]}
REL = 1e-12


def se_mean(n):           # sqrt(P) * s.e.(mean) = 1.2973 for the README model
    return 1.2973 / math.sqrt(n)


# ---------------------------------------------------------------------------------------------------
# replay gate
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,kind,per", [("gaussian_unknown_mean", "real", 1), ("gaussian_unknown_mean_mu", "real", 1),
                                             ("linear_gaussian_1d", "real", 32), ("hmm", "int", 64)])
def test_replay_of_reference_traces(engine, oracle, tmp_path, model, kind, per):
    prefix = str(tmp_path / "ref")
    n = 20_000
    oracle.run(model, OBS[model], n, prefix, how="fast", seed=17)
    _, values, logw_file = oracle.parse_records(f"{prefix}.{kind}", kind, per, n)
    rows = np.ascontiguousarray(values.T)
    got = engine.replay(model, OBS[model], real_rows=rows if kind == "real" else None, int_rows=rows if kind == "int" else None)
    # (a) against the log-weights in the reference-format file (16 significant digits of text)
    np.testing.assert_allclose(got, logw_file, rtol=REL)
    # (b) against the oracle replaying the same parsed values in full double precision
    exact = oracle.replay_logw(model, OBS[model], values)
    np.testing.assert_allclose(got, exact, rtol=REL)
    assert np.abs(got - exact).max() <= 1e-13 * np.abs(exact).max()


def test_replay_special_values(engine, oracle):
    # x = +-inf observations and out-of-range states give -inf exactly like the reference's logpdf
    lw = engine.replay("gaussian_unknown_mean", [math.inf, 4.0], real_rows=np.array([[0.5, 1.0]]))
    assert np.isneginf(lw).all()
    vals = np.array([[0.3], [2.0], [-1.5]])
    np.testing.assert_allclose(engine.replay("gaussian_unknown_mean", [3.0, 4.0], real_rows=vals.T),
                               oracle.replay_logw("gaussian_unknown_mean", [3.0, 4.0], vals), rtol=REL)
    assert abs(engine.replay("gaussian_unknown_mean", [3.0, 4.0], real_rows=np.array([[2.0]]))[0] - RM["log_w_at_mu_2"]) < 1e-14


# ---------------------------------------------------------------------------------------------------
# posterior moments: analytic + reference CPU SIS
# ---------------------------------------------------------------------------------------------------
def test_readme_model_10k_vs_analytic_and_cpu_sis(engine, oracle, tmp_path):
    n = 10_000                                                   # BASELINE.json configs[0]
    st = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    assert st["n_particles"] == n and st["n_real"] == 1 and st["n_int"] == 0
    assert abs(st["real_mean"][0] - RM["post_mean"]) < 4 * se_mean(n)
    assert abs(st["real_var"][0] - RM["post_var"]) < 4 * 2.2 / math.sqrt(n)
    assert abs(st["log_evidence"] - RM["log_evidence"]) < 4 * 1.0 / math.sqrt(n)
    assert abs(st["ess"] / n - RM["ess_fraction"]) < 0.03
    prefix = str(tmp_path / "cpu")
    oracle.run("gaussian_unknown_mean", [3.0, 4.0], n, prefix, how="faithful", seed=99)
    _, _, mean, var = oracle.stats_real(prefix)
    assert abs(st["real_mean"][0] - mean[0]) < 4 * math.sqrt(2) * se_mean(n)
    assert abs(st["real_var"][0] - var[0]) < 4 * math.sqrt(2) * 2.2 / math.sqrt(n)


def test_readme_model_large(engine):
    n = 1 << 28                                                  # 2.7e8 particles, a few ms on a B200
    st = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    assert abs(st["real_mean"][0] - RM["post_mean"]) < 4 * se_mean(n)
    assert abs(st["real_var"][0] - RM["post_var"]) < 4 * 2.2 / math.sqrt(n)
    assert abs(st["log_evidence"] - RM["log_evidence"]) < 4 * 1.0 / math.sqrt(n)
    assert abs(st["ess"] / n - RM["ess_fraction"]) < 1e-3
    assert st["n_neg_inf"] == 0 and st["passes"] == 1


def test_models_hpp_variant(engine):
    n = 1 << 24
    mv = G["models_hpp_variant"]
    st = engine.run("gaussian_unknown_mean_mu", mv["obs"], n)
    assert abs(st["real_mean"][0] - mv["post_mean"]) < 4 * 2.0 / math.sqrt(n)
    assert abs(st["real_var"][0] - mv["post_var"]) < 4 * 3.0 / math.sqrt(n)
    assert abs(st["log_evidence"] - mv["log_evidence"]) < 4 * 1.5 / math.sqrt(n)
    st = engine.run("gaussian_unknown_mean_mu", mv["thesis_obs"], n)       # thesis p.85: N(7.25, 5/6)
    assert abs(st["real_mean"][0] - mv["thesis_mean"]) < 0.02 and abs(st["real_var"][0] - mv["thesis_var"]) < 0.02


def ess_tolerance(st, scale):
    return 5.0 * scale / math.sqrt(max(st["ess"], 1.0))


def test_linear_gaussian_vs_kalman(engine):
    obs = G["obs_linear_gaussian_32"][:8]
    n = 1 << 24
    st = engine.run("linear_gaussian_1d", obs, n)
    ms, vs, le = analytic.kalman_smoother(obs)
    assert st["n_real"] == 8
    tol = ess_tolerance(st, 1.0)
    assert tol < 0.05
    np.testing.assert_allclose(st["real_mean"], ms, atol=tol)
    np.testing.assert_allclose(st["real_var"], vs, atol=3 * tol)
    assert abs(st["log_evidence"] - le) < 3 * tol


def test_hmm_vs_forward_backward(engine):
    obs = G["obs_hmm_64"][:12]
    n = 1 << 24
    st = engine.run("hmm", obs, n)
    post, le = analytic.hmm_forward_backward(obs)
    assert st["n_int"] == 12 and st["int_lo"] == 0 and st["int_bins"] == 3
    tol = ess_tolerance(st, 1.0)
    assert tol < 0.02
    np.testing.assert_allclose(st["int_prob"], post, atol=tol)
    assert (st["int_map"] == post.argmax(1)).all()
    assert abs(st["log_evidence"] - le) < 3 * tol
    np.testing.assert_allclose(st["int_prob"].sum(1), 1.0, rtol=1e-12)


# ---------------------------------------------------------------------------------------------------
# the on-device reduction IS the reference's StatsPrinter arithmetic: run it on identical records
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,n_obs", [("gaussian_unknown_mean", 2), ("linear_gaussian_1d", 32), ("hmm", 64)])
def test_device_estimators_equal_stats_printer_on_same_records(engine, oracle, tmp_path, model, n_obs):
    prefix = str(tmp_path / "gpu")
    n = 3 * capi.CHUNK + 1234                      # several chunks, ragged tail
    obs = OBS[model][:n_obs]
    st = engine.infer_to_files(model, obs, n, prefix)
    if st["n_real"]:
        ids, ks, mean, var = oracle.stats_real(prefix, max_rows=128)
        assert ks.tolist() == list(range(st["n_real"]))
        # text carries 16 significant digits, and with 32 steps the effective sample is tiny: compare
        # relative to the spread of the values
        np.testing.assert_allclose(st["real_mean"], mean, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st["real_var"], var, rtol=1e-7, atol=1e-9)
    if st["n_int"]:
        ids, ks, prob, mp, npts = oracle.stats_int(prefix, int(st["int_lo"]), int(st["int_bins"]), max_rows=128)
        np.testing.assert_allclose(st["int_prob"], prob, rtol=1e-9, atol=1e-12)
        assert (st["int_map"] == mp).all() and (npts == n).all()


def test_reduce_records_is_stats_printer(engine, oracle, tmp_path):
    """cpprob_sis_reduce_records on records written by the (restated) reference."""
    prefix = str(tmp_path / "ref")
    n = 50_000
    obs = OBS["linear_gaussian_1d"][:6]
    oracle.run("linear_gaussian_1d", obs, n, prefix, seed=4)
    _, values, logw = oracle.parse_records(prefix + ".real", "real", 6, n)
    st = engine.reduce_records(logw, real_rows=values.T)
    _, _, mean, var = oracle.stats_real(prefix)
    np.testing.assert_allclose(st["real_mean"], mean, rtol=1e-11)
    np.testing.assert_allclose(st["real_var"], var, rtol=1e-9)
    obs = OBS["hmm"][:10]
    prefix = str(tmp_path / "refh")
    oracle.run("hmm", obs, n, prefix, seed=5)
    _, values, logw = oracle.parse_records(prefix + ".int", "int", 10, n)
    st = engine.reduce_records(logw, int_rows=values.T)
    _, _, prob, mp, _ = oracle.stats_int(prefix, int(st["int_lo"]), int(st["int_bins"]))
    np.testing.assert_allclose(st["int_prob"], prob, rtol=1e-11, atol=1e-15)
    assert (st["int_map"] == mp).all()


# ---------------------------------------------------------------------------------------------------
# trace emission: rows == what the estimators saw; fused and row paths agree
# ---------------------------------------------------------------------------------------------------
def test_emitted_trace_consistency(engine):
    n = 2 * capi.CHUNK + 777
    out = engine.run("gaussian_unknown_mean", [3.0, 4.0], n, collect=True)
    mu, lw = out["real_rows"][0], out["log_w"]
    assert mu.shape == (n,) and lw.shape == (n,)
    # log_w is the reference's accumulation: 0.0 + logpdf(3) + logpdf(4)
    ref = engine.replay("gaussian_unknown_mean", [3.0, 4.0], real_rows=mu[None, :])
    assert (ref == lw).all()                                      # same device code path -> bit-exact
    w = np.exp(lw - lw.max())
    np.testing.assert_allclose(out["real_mean"][0], (w * mu).sum() / w.sum(), rtol=1e-12)
    np.testing.assert_allclose(out["log_sum_exp"], lw.max() + math.log(w.sum()), rtol=1e-13)
    np.testing.assert_allclose(out["ess"], w.sum() ** 2 / (w * w).sum(), rtol=1e-12)
    assert out["max_log_w"] == lw.max()
    # prior draws: mu ~ N(1, 1.5)
    assert abs(mu.mean() - 1.0) < 5 * 1.5 / math.sqrt(n) and abs(mu.std() - 1.5) < 0.02
    fused = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    rows = engine.run("gaussian_unknown_mean", [3.0, 4.0], n, force_rows=True)
    for k in ("real_mean", "real_var"):
        np.testing.assert_allclose(fused[k], out[k], rtol=1e-12)
        np.testing.assert_allclose(rows[k], out[k], rtol=0, atol=0)
    assert fused["max_log_w"] == out["max_log_w"] and fused["ess"] == pytest.approx(out["ess"], rel=1e-12)


def test_particle_streams_are_keyed_by_global_index(engine):
    """The sample of particle p does not depend on how many particles are run."""
    a = engine.run("gaussian_unknown_mean", [3.0, 4.0], 1000, collect=True)["real_rows"][0]
    b = engine.run("gaussian_unknown_mean", [3.0, 4.0], capi.CHUNK + 1000, collect=True)["real_rows"][0]
    c = engine.run("linear_gaussian_1d", G["obs_linear_gaussian_32"][:3], 600, collect=True)["real_rows"]
    d = engine.run("linear_gaussian_1d", G["obs_linear_gaussian_32"][:3], 5000, collect=True)["real_rows"]
    # the tile layout pairs p with p+256 only when both exist; compare full tiles
    assert (a[:512] == b[:512]).all()
    assert (c[:, :512] == d[:, :512]).all()
    assert len(np.unique(b)) == b.size


# ---------------------------------------------------------------------------------------------------
# determinism: grid size, repeated runs, sharding
# ---------------------------------------------------------------------------------------------------
def test_bitwise_reproducible_for_any_grid_and_shard_count(engine):
    import ctypes
    from cpprob_b200 import Engine
    n = 37 * capi.CHUNK + 4321
    base = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    again = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    assert (base["sums"] == again["sums"]).all()
    with Engine(seed=0x5EED, blocks_per_sm=1) as small:
        other = small.run("gaussian_unknown_mean", [3.0, 4.0], n)
    assert (base["sums"] == other["sums"]).all()
    with Engine(seed=0x5EED + 1) as reseeded:
        assert (reseeded.run("gaussian_unknown_mean", [3.0, 4.0], n)["sums"] != base["sums"]).any()
    # emulate 1, 2, 3 and 8 ranks on one GPU: concatenate the shard partials in rank order and merge
    import torch
    for world in (1, 2, 3, 8):
        parts, m_ref, n_cols = [], None, None
        for r in range(world):
            p = engine.run_shard("gaussian_unknown_mean", [3.0, 4.0], n, r, world)
            m_ref, n_cols = p.m_ref, p.n_cols
            t = torch.empty((p.n_chunks_local, p.n_cols), dtype=torch.float64, device="cuda")
            if p.n_chunks_local:
                ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(p.device_ptr),
                                                       ctypes.c_size_t(t.numel() * 8), 3)
            parts.append(t)
        g = torch.cat(parts).contiguous()
        torch.cuda.synchronize()
        st, rebase = engine.merge("gaussian_unknown_mean", [3.0, 4.0], g.data_ptr(), g.shape[0], n_cols, m_ref, n)
        assert not rebase
        assert (st["sums"] == base["sums"]).all(), world
        assert st["real_mean"][0] == base["real_mean"][0]


@pytest.mark.parametrize("model,obs", [("gaussian_unknown_mean", [3.0, 4.0]), ("hmm", None)])
def test_super_chunk_rows_are_world_size_invariant(engine, model, obs):
    """Beyond 4096 chunks a rank reduces super-chunks (2^k chunks) to one row each before the gather
    (cpprob_sis_plan_rows): the exchange stays <= 4096 rows and the merged sums are still bit-identical for 1, 2, 3
    and 8 ranks, on the fused path and on the row path (8 kernel rows per chunk)."""
    import ctypes
    import torch
    obs = G["obs_hmm_64"][:3] if obs is None else obs
    n = 5000 * capi.CHUNK + 777                      # 5001 chunks -> super-chunks of 2, 2501 rows
    base = engine.run(model, obs, n)
    for world in (1, 2, 3, 8):
        parts, m_ref, n_cols, total_rows = [], None, None, None
        for r in range(world):
            p = engine.run_shard(model, obs, n, r, world)
            m_ref, n_cols, total_rows = p.m_ref, p.n_cols, p.n_chunks_total
            assert (p.chunk_first, p.n_chunks_local, p.n_chunks_total) == capi.plan_rows(n, r, world, p.rows_per_chunk)
            t = torch.empty((p.n_chunks_local, p.n_cols), dtype=torch.float64, device="cuda")
            if p.n_chunks_local:
                ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(p.device_ptr),
                                                       ctypes.c_size_t(t.numel() * 8), 3)
            parts.append(t)
        g = torch.cat(parts).contiguous()
        torch.cuda.synchronize()
        assert g.shape[0] == total_rows == 2501
        st, rebase = engine.merge(model, obs, g.data_ptr(), g.shape[0], n_cols, m_ref, n)
        assert not rebase
        assert (st["sums"] == base["sums"]).all(), world
        # the raw all-gather layout (segments padded to the largest shard, padding never initialised) merges to the same bits
        m = max(t.shape[0] for t in parts)
        padded = torch.full((world, m, n_cols), float("nan"), dtype=torch.float64, device="cuda")
        for r, t in enumerate(parts):
            padded[r, :t.shape[0]] = t
        torch.cuda.synchronize()
        st2, rebase2 = engine.merge_padded(model, obs, padded.data_ptr(), world, m, p.rows_per_chunk, n_cols, m_ref, n)
        assert not rebase2 and (st2["sums"] == base["sums"]).all(), world
    assert base["n_particles"] == n


# ---------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 511, 512, 513, capi.CHUNK - 1, capi.CHUNK, capi.CHUNK + 1])
def test_ragged_particle_counts(engine, n):
    out = engine.run("gaussian_unknown_mean", [3.0, 4.0], n, collect=True)
    mu, lw = out["real_rows"][0], out["log_w"]
    assert mu.size == n
    w = np.exp(lw - lw.max())
    np.testing.assert_allclose(out["real_mean"][0], (w * mu).sum() / w.sum(), rtol=1e-12)
    fused = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    np.testing.assert_allclose(fused["real_mean"][0], out["real_mean"][0], rtol=1e-12)
    st = engine.run("hmm", G["obs_hmm_64"][:5], n)
    assert st["n_particles"] == n and abs(st["int_prob"].sum() - 5.0) < 1e-9


def test_all_weights_minus_infinity(engine):
    n = 5000
    st = engine.run("gaussian_unknown_mean", [math.inf, 4.0], n)          # logpdf(|x| = inf) = -inf
    assert st["n_neg_inf"] == n and st["max_log_w"] == -math.inf
    assert math.isnan(st["real_mean"][0])                                  # the reference yields NaN too (0/0)


@pytest.mark.parametrize("x2", [4.0e3, 1.0e6, 3.0e6])
def test_widely_spread_log_weights(engine, x2):
    """Log-weights spread over up to 1e12 below the maximum: the fused kernel's table-assisted exp only holds for
    arguments >= -707 and its integer exponent wraps for |argument| > 3.6e5, so the units must be caught on the
    argument itself and redone by the careful pass; the row path (plain exp_weight) is the cross-check."""
    n = 2_000_000
    a = engine.run("gaussian_unknown_mean", [3.0, x2], n)
    b = engine.run("gaussian_unknown_mean", [3.0, x2], n, force_rows=True)
    assert a["n_nan"] == 0 and b["n_nan"] == 0
    np.testing.assert_allclose(a["real_mean"], b["real_mean"], rtol=1e-12)
    np.testing.assert_allclose(a["log_evidence"], b["log_evidence"], rtol=1e-12)
    np.testing.assert_allclose(a["ess"], b["ess"], rtol=1e-9)
    assert a["max_log_w"] == b["max_log_w"]


def test_rebase_when_reference_is_far_off(engine):
    n = 4 * capi.CHUNK
    base = engine.run("gaussian_unknown_mean", [3.0, 4.0], n)
    p = engine.run_shard("gaussian_unknown_mean", [3.0, 4.0], n, 0, 1, m_ref=base["max_log_w"] + 2000.0)
    st, rebase = engine.merge("gaussian_unknown_mean", [3.0, 4.0], p.device_ptr, p.n_chunks_total, p.n_cols, p.m_ref, n)
    assert rebase and st["max_log_w"] == base["max_log_w"]
    p = engine.run_shard("gaussian_unknown_mean", [3.0, 4.0], n, 0, 1, m_ref=st["max_log_w"])
    st2, rebase2 = engine.merge("gaussian_unknown_mean", [3.0, 4.0], p.device_ptr, p.n_chunks_total, p.n_cols, p.m_ref, n)
    assert not rebase2
    np.testing.assert_allclose(st2["real_mean"], base["real_mean"], rtol=1e-12)
    np.testing.assert_allclose(st2["log_evidence"], base["log_evidence"], rtol=1e-12)


def test_error_behaviour(engine):
    with pytest.raises(capi.SisError) as ei:
        engine.run("gaussian_unknown_mean", [3.0], 100)
    assert ei.value.code == -1 and "takes 2 observations" in str(ei.value)
    with pytest.raises(capi.SisError) as ei:
        engine.run("nope", [3.0], 100)
    assert ei.value.code == -3
    with pytest.raises(capi.SisError):
        engine.run("gaussian_unknown_mean", [3.0, 4.0], 0)


def test_structure_probe(engine):
    d = engine.describe("gaussian_unknown_mean", [3.0, 4.0])
    assert d["ids"] == ["Mean"] and d["slots"] == [(0, 0, 0, 0)] and d["n_samples"] == 1
    d = engine.describe("gaussian_unknown_mean_mu", [3.0, 4.0])
    assert d["ids"] == ["Mu"]
    d = engine.describe("linear_gaussian_1d", G["obs_linear_gaussian_32"])
    assert d["ids"] == ["State"] and d["n_real"] == 32 and [s[2] for s in d["slots"]] == list(range(32))
    d = engine.describe("hmm", G["obs_hmm_64"])
    assert d["ids"] == ["State"] and d["n_int"] == 64 and d["n_real"] == 0 and d["n_samples"] == 64


def test_single_process_multi_gpu_is_bit_identical():
    """cpprob_sis_run_multi: shards on several GPUs of one process, peer-copied partials, merge on GPU 0."""
    import torch
    from cpprob_b200 import Engine
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 61 * capi.CHUNK + 99
    with Engine(device=0, seed=77) as ref:
        base = ref.run("gaussian_unknown_mean", [3.0, 4.0], n)
        base_hmm = ref.run("hmm", G["obs_hmm_64"][:9], n // 8)
    engines = [Engine(device=d, seed=77) for d in range(min(n_dev, 8))]
    try:
        st = capi.run_multi(engines, "gaussian_unknown_mean", [3.0, 4.0], n)
        assert (st["sums"] == base["sums"]).all() and st["real_mean"][0] == base["real_mean"][0]
        st = capi.run_multi(engines, "hmm", G["obs_hmm_64"][:9], n // 8)
        assert (st["sums"] == base_hmm["sums"]).all()
        # beyond 4096 chunks every GPU hands on super-chunk rows (cpprob_sis_plan_rows)
        n_big = 5000 * capi.CHUNK + 777
        with Engine(device=0, seed=77) as ref:
            base_big = ref.run("gaussian_unknown_mean", [3.0, 4.0], n_big)
        st = capi.run_multi(engines, "gaussian_unknown_mean", [3.0, 4.0], n_big)
        assert (st["sums"] == base_big["sums"]).all()
    finally:
        for e in engines:
            e.close()


# ---------------------------------------------------------------------------------------------------
# C5: hmm with the 1000-step observation sequence (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------------
OBS_1000 = G["obs_hmm_1000"]


def test_c5_replay_of_reference_traces(engine, oracle, tmp_path):
    prefix = str(tmp_path / "ref1000")
    n = 1500
    oracle.run("hmm", OBS_1000, n, prefix, how="fast", seed=41)
    _, values, logw_file = oracle.parse_records(prefix + ".int", "int", 1000, n)
    got = engine.replay("hmm", OBS_1000, int_rows=np.ascontiguousarray(values.T))
    np.testing.assert_allclose(got, logw_file, rtol=REL)
    exact = oracle.replay_logw("hmm", OBS_1000, values)
    np.testing.assert_allclose(got, exact, rtol=REL)
    assert np.abs(got - exact).max() <= 1e-13 * np.abs(exact).max()


def test_c5_gpu_text_equals_host_text(tmp_path, monkeypatch):
    import os
    from cpprob_b200 import Engine
    n = capi.CHUNK + 333                           # two batches of 6 KB lines
    with Engine(seed=0xC5, max_batch=capi.CHUNK) as e:
        monkeypatch.delenv("CPPROB_SIS_TEXT", raising=False)
        e.infer_to_files("hmm", OBS_1000, n, str(tmp_path / "gpu"))
        monkeypatch.setenv("CPPROB_SIS_TEXT", "host")
        e.infer_to_files("hmm", OBS_1000, n, str(tmp_path / "host"))
    a = open(tmp_path / "gpu.int", "rb").read()
    assert a == open(tmp_path / "host.int", "rb").read()
    lines = a.splitlines()
    assert len(lines) == n and all(l.count(b"(0 ") == 1000 for l in lines[:20])
    assert not os.path.exists(tmp_path / "gpu.real") and open(tmp_path / "gpu.ids").read() == "State\n"


def test_c5_prefix_vs_forward_backward(engine):
    obs = OBS_1000[:12]
    n = 1 << 24
    st = engine.run("hmm", obs, n)
    post, le = analytic.hmm_forward_backward(obs)
    tol = ess_tolerance(st, 1.0)
    assert tol < 0.03
    np.testing.assert_allclose(st["int_prob"], post, atol=tol)
    assert abs(st["log_evidence"] - le) < 3 * tol


def test_c5_all_addresses_have_estimators(engine):
    """1000 (id, k) keys: every one gets a histogram that sums to 1, the first steps agree with forward-backward on the
    WHOLE sequence (smoothing marginals) as far as the collapsed effective sample allows, and the emitting and the
    estimator-only path give the same bits."""
    n = 1 << 20
    st = engine.run("hmm", OBS_1000, n)
    assert st["n_int"] == 1000 and st["int_bins"] == 3 and st["int_lo"] == 0 and st["n_particles"] == n
    assert st["int_prob"].shape == (1000, 3)
    np.testing.assert_allclose(st["int_prob"].sum(1), 1.0, rtol=1e-12)
    assert ((st["int_map"] >= 0) & (st["int_map"] <= 2)).all()
    assert np.isfinite(st["log_evidence"]) and st["n_neg_inf"] == 0 and 1.0 <= st["ess"] <= n
    _, le = analytic.hmm_forward_backward(OBS_1000)
    # importance sampling from the prior over 1000 steps: the estimate of the evidence is dominated by the best particle
    # and biased low; it cannot exceed the truth by more than noise
    assert st["log_evidence"] < le + 5.0
    rows = engine.run("hmm", OBS_1000, n, force_rows=True)
    assert (rows["sums"] == st["sums"]).all()
