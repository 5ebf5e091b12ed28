#!/usr/bin/env python3
"""Where the fixed cost of one inference goes (README model) at the per-GPU share of the strong-scaling point.
Run plain for CUDA-event / wall times, or under `ncu --metrics gpu__time_duration.sum` for the serialised kernel list.
usage: python tools/strong_breakdown.py [particles] [reps] [sha]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpprob_b200 import Engine  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with Engine(seed=0x5EED) as e:
    best_w, best_d = 1e9, 1e9
    for i in range(reps):
        t0 = time.perf_counter()
        st = e.run("gaussian_unknown_mean", [3.0, 4.0], n)
        w = time.perf_counter() - t0
        if i >= min(5, reps - 1):
            best_w, best_d = min(best_w, w), min(best_d, st["device_ms"])
    print(f"n={n} wall {best_w * 1e3:.4f} ms device {best_d:.4f} ms launches {st['kernel_launches']}")
    if len(sys.argv) > 3:                      # fingerprint of the merged sums of the 2^30-particle run (bench.py's sums_sha)
        import hashlib
        print("sums_sha", hashlib.sha256(e.run("gaussian_unknown_mean", [3.0, 4.0], 1 << 30)["sums"].tobytes()).hexdigest()[:16])
