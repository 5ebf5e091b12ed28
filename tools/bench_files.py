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
CASES = (("README model", "gaussian_unknown_mean", [3.0, 4.0], n), ("hmm<64>", "hmm", g["obs_hmm_64"], n // 8),
         ("linear_gaussian<32>", "linear_gaussian_1d", g["obs_linear_gaussian_32"], n // 16),
         ("hmm<1000>", "hmm", g["obs_hmm_1000"] if "obs_hmm_1000" in g else list(g["obs_hmm_64"]) * 16, n // 128))
for mode in ("gpu", "host"):
    if mode == "host":
        os.environ["CPPROB_SIS_TEXT"] = "host"
    else:
        os.environ.pop("CPPROB_SIS_TEXT", None)
    with Engine(seed=1) as e, tempfile.TemporaryDirectory(dir=d) as tmp:
        for name, model, obs, m in CASES:        # first round: warm-up (pinned allocations, page cache)
            e.infer_to_files(model, obs, m, os.path.join(tmp, "warm_" + model + str(len(obs))))
        for f in os.listdir(tmp):
            os.remove(os.path.join(tmp, f))
        for name, model, obs, m in CASES:
            prefix = os.path.join(tmp, model + str(len(obs)))
            t0 = time.perf_counter()
            st = e.infer_to_files(model, obs, m, prefix)
            wall = time.perf_counter() - t0
            size = sum(os.path.getsize(prefix + ext) for ext in (".real", ".int") if os.path.exists(prefix + ext))
            line = (f"[{mode} text] {name}: {m} records in {wall:.3f} s = {m / wall:.3e} records/s, {size / wall / 1e9:.2f} GB/s of text "
                    f"({size / m:.0f} B/record), particle kernels {st['device_ms']:.1f} ms")
            if mode == "gpu":
                ts = e.text_stage_stats()
                line += (f"; text kernels {ts['kernel_ms']:.1f} ms ({ts['bytes'] / max(ts['kernel_ms'], 1e-9) / 1e6:.0f} GB/s), "
                         f"D2H {ts['copy_ms']:.1f} ms ({ts['bytes'] / max(ts['copy_ms'], 1e-9) / 1e6:.1f} GB/s), file writes {ts['write_s']:.3f} s "
                         f"({ts['bytes'] / max(ts['write_s'], 1e-9) / 1e9:.1f} GB/s)")
            print(line, flush=True)
            for ext in (".real", ".int", ".ids", ".stats"):
                if os.path.exists(prefix + ext):
                    os.remove(prefix + ext)
