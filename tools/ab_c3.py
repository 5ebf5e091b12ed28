#!/usr/bin/env python3
"""A/B timing of C3 (linear-Gaussian, 32 steps): staged path at 1e8 particles and row path at 2^24, best of 5 after a warm-up.
usage: CPPROB_SIS_LIB=<variant.so> python tools/ab_c3.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import analytic  # noqa: E402
from cpprob_b200 import Engine  # noqa: E402

g = analytic.golden()
obs = g["obs_linear_gaussian_32"]
with Engine(seed=0x5EED) as e:
    for label, n, kw in (("staged", 100_000_000, {}), ("rows", 1 << 24, {"force_rows": True})):
        best = min(e.run("linear_gaussian_1d", obs, n, **kw)["device_ms"] for _ in range(6))
        print(f"{os.environ.get('CPPROB_SIS_LIB', 'shipped')}: C3 {label} {n} particles {best:.3f} ms = {n / best / 1e3:.4g} particles/s")
