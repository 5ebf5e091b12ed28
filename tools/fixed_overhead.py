#!/usr/bin/env python3
"""Fixed cost of one inference call (README model): wall time and device time of cpprob_sis_run for a range of particle
counts, best of 20.  The intercept is what strong scaling pays per GPU count.  usage: python tools/fixed_overhead.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpprob_b200 import Engine  # noqa: E402

with Engine(seed=0x5EED) as e:
    for n in (10_000, 1_000_000, 31_250_000, 125_000_000, 250_000_000, 1_000_000_000):
        best_w, best_d = 1e9, 1e9
        for i in range(25):
            t0 = time.perf_counter()
            st = e.run("gaussian_unknown_mean", [3.0, 4.0], n)
            w = time.perf_counter() - t0
            if i >= 5:
                best_w, best_d = min(best_w, w), min(best_d, st["device_ms"])
        print(f"n={n:>12d}  wall {best_w * 1e3:8.3f} ms  device {best_d:8.3f} ms  launches {st['kernel_launches']}  ideal {n / 3.24e11 * 1e3:8.3f} ms")
