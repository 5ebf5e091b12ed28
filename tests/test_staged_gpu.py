"""The staged path (k_sis_staged: traces in per-warp shared-memory staging areas, estimators formed in the same kernel,
no trace row in HBM) against the row path (k_sis_rows + k_rows_moments / k_rows_hist) and against records.

Both paths form every per-(address, k) sum in the same canonical order with the same round functions
(cpprob_b200/csrc/staged_kernels.cuh), so their merged sums must be equal bit for bit, for any particle count, any
schedule of the work units and any number of ranks."""
import ctypes

import numpy as np
import pytest

import analytic
from cpprob_b200 import capi

pytestmark = pytest.mark.gpu
G = analytic.golden()

CASES = [("hmm", G["obs_hmm_1000"]), ("hmm", G["obs_hmm_1000"][:130]), ("linear_gaussian_1d", G["obs_linear_gaussian_32"]), ("linear_gaussian_1d", G["obs_linear_gaussian_32"][:5]),
         ("hmm", G["obs_hmm_64"]), ("hmm", G["obs_hmm_64"][:7]), ("hmm", G["obs_hmm_64"][:1]),
         ("linear_gaussian_1d", (G["obs_linear_gaussian_32"] * 2)[:45])]         # 45 real rows: two row groups per lane


@pytest.mark.parametrize("model,obs", CASES, ids=[f"{m}{len(o)}" for m, o in CASES])
@pytest.mark.parametrize("n", [1, 31, 33, 256, 511, 513, 4095, 4097, 3 * capi.CHUNK + 1234])
def test_staged_sums_equal_row_path_bits(engine, model, obs, n):
    a = engine.run(model, obs, n)
    b = engine.run(model, obs, n, force_rows=True)
    assert a["path"] == "staged" and b["path"] == "rows"
    assert a["n_cols"] == b["n_cols"]
    assert (a["sums"] == b["sums"]).all(), np.nonzero(a["sums"] != b["sums"])
    for k in ("real_mean", "real_var", "int_prob", "int_map", "log_sum_exp", "ess", "max_log_w", "n_neg_inf"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_path_selection(engine):
    assert engine.run("gaussian_unknown_mean", [3.0, 4.0], 1000)["path"] == "fused"          # 1 real predict: registers
    assert engine.run("linear_regression", [0.0, 1.0, 1.0, 3.0], 1000)["path"] == "fused"     # 2 real predicts
    assert engine.run("linear_gaussian_1d", G["obs_linear_gaussian_32"], 1000)["path"] == "staged"
    assert engine.run("hmm", G["obs_hmm_64"], 1000)["path"] == "staged"
    assert engine.run("hmm", G["obs_hmm_64"], 1000, force_rows=True)["path"] == "rows"
    assert engine.run("hmm", G["obs_hmm_64"], 1000, collect=True)["path"] == "rows"           # emitting runs need the rows
    # 1000 int predicts of a model that declares 3 states: four states per staged byte, 8 KB per warp
    st = engine.run("hmm", G["obs_hmm_1000"], 2000)
    assert st["path"] == "staged" and st["n_int"] == 1000 and st["int_lo"] == 0 and st["int_bins"] == 3
    # far longer traces fall back to rows (neither the staging areas nor the model's table fit)
    long_obs = (G["obs_hmm_1000"] * 30)[:30000]
    assert engine.run("hmm", long_obs, 300)["path"] == "rows"


def test_staged_matches_records(engine):
    """The estimators of the staged path are those of the emitted records (numpy, independent summation)."""
    n = 2 * capi.CHUNK + 99
    obs = G["obs_linear_gaussian_32"][:9]
    st = engine.run("linear_gaussian_1d", obs, n)
    rec = engine.run("linear_gaussian_1d", obs, n, collect=True)
    w = np.exp(rec["log_w"] - rec["log_w"].max())
    for k in range(9):
        x = rec["real_rows"][k]
        m = (w * x).sum() / w.sum()
        assert abs(st["real_mean"][k] - m) <= 1e-11 * max(1.0, abs(m))
        assert abs(st["real_var"][k] - ((w * x * x).sum() / w.sum() - m * m)) <= 1e-10
    obs = G["obs_hmm_64"][:20]
    st = engine.run("hmm", obs, n)
    rec = engine.run("hmm", obs, n, collect=True)
    w = np.exp(rec["log_w"] - rec["log_w"].max())
    for k in range(20):
        for v in range(3):
            assert abs(st["int_prob"][k, v] - w[rec["int_rows"][k] == v].sum() / w.sum()) <= 1e-12
    assert st["ess"] == pytest.approx(w.sum() ** 2 / (w * w).sum(), rel=1e-12)


@pytest.mark.parametrize("model,obs", [("linear_gaussian_1d", G["obs_linear_gaussian_32"][:6]), ("hmm", G["obs_hmm_64"][:9])])
def test_staged_reproducible_and_rank_count_invariant(engine, model, obs):
    import torch
    n = 21 * capi.CHUNK + 4321
    base = engine.run(model, obs, n)
    assert base["path"] == "staged"
    assert (engine.run(model, obs, n)["sums"] == base["sums"]).all()        # any schedule of the (sub-chunk, slot) units
    for world in (1, 2, 3, 8):
        parts, m_ref, n_cols = [], None, None
        for r in range(world):
            p = engine.run_shard(model, obs, n, r, world)
            m_ref, n_cols = p.m_ref, p.n_cols
            assert (p.chunk_first, p.n_chunks_local, p.n_chunks_total) == capi.plan_rows(n, r, world, p.rows_per_chunk)
            t = torch.empty((p.n_chunks_local, p.n_cols), dtype=torch.float64, device="cuda")
            if p.n_chunks_local:
                ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(p.device_ptr),
                                                       ctypes.c_size_t(t.numel() * 8), 3)
            parts.append(t)
        g = torch.cat(parts).contiguous()
        torch.cuda.synchronize()
        st, rebase = engine.merge(model, obs, g.data_ptr(), g.shape[0], n_cols, m_ref, n)
        assert not rebase and (st["sums"] == base["sums"]).all(), world


def test_staged_window_widening(engine):
    """An int value outside the pilot's histogram window makes the run repeat with a wider one (as on the row path)."""
    # all_distr draws a Poisson(0.8): the pilot's 4096 particles rarely see its largest values
    n = 1 << 22
    a = engine.run("all_distr", [0.0, 0.0], n)
    b = engine.run("all_distr", [0.0, 0.0], n, force_rows=True)
    assert a["int_bins"] == b["int_bins"] and a["int_lo"] == b["int_lo"]
    np.testing.assert_allclose(a["int_prob"], b["int_prob"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(a["real_mean"], b["real_mean"], rtol=1e-12)


def test_rebase_keeps_ess_finite(engine):
    """Log-weights far above the pilot's maximum: sum w^2 must not overflow (the re-base limit is 300, not 600)."""
    # an override 400 below the true maximum puts every weight near e^400: w^2 would overflow without the re-base pass
    p = engine.run_shard("gaussian_unknown_mean", [3.0, 4.0], 100_000, 0, 1, m_ref=-404.0)
    st, rebase = engine.merge("gaussian_unknown_mean", [3.0, 4.0], p.device_ptr, p.n_chunks_total, p.n_cols, p.m_ref, 100_000)
    assert rebase, "weights e^400 above the reference must trigger the re-base pass"
    p = engine.run_shard("gaussian_unknown_mean", [3.0, 4.0], 100_000, 0, 1, m_ref=st["max_log_w"])
    st, rebase = engine.merge("gaussian_unknown_mean", [3.0, 4.0], p.device_ptr, p.n_chunks_total, p.n_cols, p.m_ref, 100_000)
    assert not rebase and np.isfinite(st["ess"]) and 0.4 < st["ess"] / 100_000 < 0.6
