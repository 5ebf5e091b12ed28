#!/usr/bin/env python3
"""Statistical calibration of the staged path with its batched draws (next_std_normal_x4, next_u32x4, table-driven hmm):
over S independent seeds the per-(address, k) posterior estimates of linear_gaussian_1d / hmm must scatter around the exact
Kalman / forward-backward posteriors like their own standard error says, and the evidence estimate (unbiased) around the exact
evidence.  usage: python tools/calibration_staged.py [n_seeds] [log2 particles]"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import analytic  # noqa: E402
from cpprob_b200 import Engine  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
g = analytic.golden()


def report(name, est, exact, evid, le):
    est = np.array(est)                       # [S, K]
    m, sd = est.mean(0), est.std(0, ddof=1)
    z = (m - exact) / (sd / math.sqrt(S))
    ratio = np.exp(np.array(evid) - le)       # Z_hat / Z: mean 1
    zr = (ratio.mean() - 1.0) / (ratio.std(ddof=1) / math.sqrt(S))
    print(f"{name}: {S} seeds x {n} particles; z of the seed-mean of each estimate against the exact posterior:")
    print("   " + " ".join(f"{v:+.2f}" for v in z))
    print(f"   max |z| {np.abs(z).max():.2f} over {z.size} estimates (expect < ~3.5), mean z {z.mean():+.2f}; evidence ratio mean "
          f"{ratio.mean():.5f} (z {zr:+.2f})")
    return float(np.abs(z).max()), float(zr)


with Engine(seed=0) as e:
    obs = g["obs_linear_gaussian_32"][:8]
    ms, vs, le = analytic.kalman_smoother(obs)
    est, evid = [], []
    for s in range(S):
        e.set_seed(5000 + s)
        st = e.run("linear_gaussian_1d", obs, n)
        assert st["path"] == "staged"
        est.append(st["real_mean"])
        evid.append(st["log_evidence"])
    a = report("linear_gaussian_1d (8 steps), posterior means", est, np.array(ms), evid, le)
    obs = g["obs_hmm_64"][:12]
    post, le = analytic.hmm_forward_backward(obs)
    est, evid = [], []
    for s in range(S):
        e.set_seed(7000 + s)
        st = e.run("hmm", obs, n)
        assert st["path"] == "staged"
        est.append(st["int_prob"].reshape(-1))
        evid.append(st["log_evidence"])
    b = report("hmm (12 steps), smoothing marginals", est, post.reshape(-1), evid, le)
ok = a[0] < 4.5 and b[0] < 4.5 and abs(a[1]) < 4 and abs(b[1]) < 4
print("calibration", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
