#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json (C3 linear-Gaussian 32 steps, C4 HMM 64 steps, C5 HMM 1000 steps with
full trace emission): device-timed particles/s and, for the row (SoA) path, the HBM bytes moved.  Not a bench.py line;
results go to profiles/ as notes.  usage: python tools/bench_configs.py [--scale 1.0]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import analytic  # noqa: E402
from cpprob_b200 import Engine  # noqa: E402


def run(e, name, model, obs, n, reps=3, **kw):
    best, st = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        st = e.run(model, obs, n, **kw)
        wall = time.perf_counter() - t0
        if best is None or st["device_ms"] < best[0]:
            best = (st["device_ms"], wall)
    bpp = 8 * st["n_real"] + 4 * st["n_int"] + 16
    row = {"config": name, "model": model, "n_obs": len(obs), "particles": n, "device_ms": best[0], "wall_s": best[1],
           "particles_per_s": n / (best[0] * 1e-3), "row_bytes_per_particle": bpp,
           "row_write_GBps_if_rows": bpp * n / (best[0] * 1e-3) / 1e9, "ess": st["ess"], "log_evidence": st["log_evidence"],
           "launches": st["kernel_launches"], "passes": st["passes"]}
    print(json.dumps(row), flush=True)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    g = analytic.golden()
    with Engine(seed=0x5EED, max_batch=int(os.environ.get("CPPROB_SIS_MAX_BATCH", "0"))) as e:
        print(json.dumps({"store_peak_GBps": e.store_peak(), "dfma_peak_tflops": e.dfma_peak()[0]}), flush=True)
        run(e, "C2 fused", "gaussian_unknown_mean", [3.0, 4.0], int(1e9 * a.scale))
        run(e, "C2 rows (forced)", "gaussian_unknown_mean", [3.0, 4.0], int(1e8 * a.scale), force_rows=True)
        run(e, "C3", "linear_gaussian_1d", g["obs_linear_gaussian_32"], int(1e8 * a.scale))
        run(e, "C4", "hmm", g["obs_hmm_64"], int(1e8 * a.scale))
        run(e, "C5 stats only", "hmm", g["obs_hmm_1000"], int(4e6 * a.scale))
        run(e, "C5 stats only, 1.6e7", "hmm", g["obs_hmm_1000"], int(1.6e7 * a.scale))
        # C5 with full trace emission to host memory (pinned D2H on the side stream; records dropped on arrival)
        t0 = time.perf_counter()
        n = int(2e6 * a.scale)
        st = e.run("hmm", g["obs_hmm_1000"], n, emit=1)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": "C5 emit to host (no text)", "particles": n, "device_ms": st["device_ms"], "wall_s": wall,
                          "d2h_GBps_wall": 4016 * n / wall / 1e9}), flush=True)


if __name__ == "__main__":
    main()
