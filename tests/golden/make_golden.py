#!/usr/bin/env python3
"""Generates tests/golden/golden.json — the committed known-answer fixtures.

Sources of truth (none of them is our own code):
  * the reference's known-answer grids for logpdf<normal> and logpdf<uniform_real>
    (/root/reference tests/cpprob/logpdf.cpp:23-35, :61-78), whose expected side is boost::math::pdf;
    Boost is absent here, so the expected values are recomputed with scipy.stats on the same grids
    (a deterministic subsample is stored; tests also sweep the full grid against scipy at run time);
  * golden values derived from the reference's formulas during the survey;
  * Random123's published Philox4x32-10 known-answer vectors;
  * the synthetic observation sequences of configs C3-C5, simulated from the models
    themselves with numpy's PCG64 seeded 20240607.
Run: python tests/golden/make_golden.py
"""
import json
import os

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))


def normal_grid():
    rows = []
    for mean in range(-10, 10):
        for std in range(1, 20):
            for f in range(1, 20):
                for i in range(-10, 10):
                    rows.append((mean / f, float(std), float(i)))
    return np.array(rows)


def uniform_grid():
    rows = []
    for a in range(-10, 10):
        for b in range(a + 1, 10):
            for f in range(1, 20):
                for i in range(-10, 10):
                    rows.append((a / f, b / f, float(i)))
    return np.array(rows)


def simulate_lg(n, rng):
    x, ys = 0.0, []
    for _ in range(n):
        x = x + rng.standard_normal()
        ys.append(x + rng.standard_normal())
    return ys


def simulate_hmm(n, rng):
    T = np.array([[0.1, 0.5, 0.4], [0.2, 0.2, 0.6], [0.15, 0.15, 0.7]])
    means = np.array([-1.0, 0.0, 1.0])
    s = rng.integers(0, 3)
    ys = []
    for t in range(n):
        if t > 0:
            s = rng.choice(3, p=T[s] / T[s].sum())
        ys.append(means[s] + rng.standard_normal())
    return ys


def main():
    g = {}
    sel = np.random.default_rng(1).choice
    ng = normal_grid()
    idx = np.sort(sel(len(ng), 1500, replace=False))
    g["logpdf_normal"] = {"mean_sigma_x": ng[idx].tolist(),
                          "expected": stats.norm.logpdf(ng[idx, 2], ng[idx, 0], ng[idx, 1]).tolist()}
    ug = uniform_grid()
    idx = np.sort(sel(len(ug), 1500, replace=False))
    with np.errstate(divide="ignore"):
        exp = stats.uniform.logpdf(ug[idx, 2], ug[idx, 0], ug[idx, 1] - ug[idx, 0])
    g["logpdf_uniform_real"] = {"a_b_x": ug[idx].tolist(), "expected": [None if np.isinf(v) else v for v in exp]}

    # SURVEY.md §8c golden values (derived from the reference's formulas)
    g["kat"] = [
        {"kind": "normal", "params": [1.0, 2.0], "x": 3.0, "expected": -2.112085713764618},
        {"kind": "normal", "params": [2.5, 2.0], "x": 4.0, "expected": -1.893335713764618},
        {"kind": "normal", "params": [0.0, 1.0], "x": 0.0, "expected": -0.9189385332046727},
        {"kind": "normal", "params": [0.7, 2 ** 0.5], "x": -2.3, "expected": -3.515512123484645},
        {"kind": "poisson", "params": [0.8], "x": 3, "expected": -3.2611901231706844},
        {"kind": "uniform_real", "params": [2.0, 9.5], "x": 5.0, "expected": -2.0149030205422647},
        {"kind": "uniform_smallint", "params": [0, 2], "x": 1, "expected": -1.0986122886681098},
        {"kind": "discrete", "params": [0.1, 0.5, 0.4], "x": 1, "expected": -0.6931471805599453},
    ]
    g["readme_model"] = {"obs": [3.0, 4.0], "log_w_at_mu_2": -3.849171427529236, "log_evidence": -4.398851446364485,
                         "ess_fraction": 0.50992, "post_mean": 2.323529411764706, "post_var": 1.0588235294117647}
    g["models_hpp_variant"] = {"obs": [3.0, 4.0], "post_mean": 3.0833333333333335, "post_var": 0.8333333333333334,
                               "log_evidence": -4.072737314916651,
                               "thesis_obs": [8.0, 9.0], "thesis_mean": 7.25, "thesis_var": 5.0 / 6.0}
    g["philox4x32_10"] = [
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
        {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
        {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
    ]
    rng = np.random.default_rng(20240607)
    g["obs_linear_gaussian_32"] = simulate_lg(32, rng)
    g["obs_hmm_64"] = simulate_hmm(64, rng)
    g["obs_hmm_1000"] = simulate_hmm(1000, rng)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=0)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
