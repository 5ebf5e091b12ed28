#!/usr/bin/env python3
"""File stage: wall time of cpprob_sis_infer_to_files for the README model and hmm<64>,
records/s and text bytes/s.  usage: python tools/bench_files.py [n]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import analytic  # noqa: E402
from cpprob_b200 import Engine  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
g = analytic.golden()
d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
with Engine(seed=1) as e, tempfile.TemporaryDirectory(dir=d) as tmp:
    e.infer_to_files("gaussian_unknown_mean", [3.0, 4.0], 100_000, os.path.join(tmp, "warm"))
    for name, model, obs, m in (("README model", "gaussian_unknown_mean", [3.0, 4.0], n), ("hmm<64>", "hmm", g["obs_hmm_64"], n // 8),
                                ("linear_gaussian<32>", "linear_gaussian_1d", g["obs_linear_gaussian_32"], n // 16)):
        prefix = os.path.join(tmp, model)
        t0 = time.perf_counter()
        st = e.infer_to_files(model, obs, m, prefix)
        wall = time.perf_counter() - t0
        size = sum(os.path.getsize(prefix + ext) for ext in (".real", ".int") if os.path.exists(prefix + ext))
        print(f"{name}: {m} records in {wall:.3f} s = {m / wall:.3e} records/s, {size / wall / 1e9:.2f} GB/s of text ({size / m:.0f} B/record), device {st['device_ms']:.1f} ms")
