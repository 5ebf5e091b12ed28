#!/usr/bin/env python3
"""One pass of each secondary configuration at a profiling-friendly size (run under ncu by tools/profile_r02.sh).
usage: python tools/configs_once.py [c3|c4|c5|c3rows|c4rows|c5rows|all] [log2 particles]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import analytic  # noqa: E402
from cpprob_b200 import Engine  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
g = analytic.golden()
with Engine(seed=0x5EED) as e:
    if which in ("c3", "all"):
        print("C3", e.run("linear_gaussian_1d", g["obs_linear_gaussian_32"], 1 << lg)["device_ms"])
    if which in ("c4", "all"):
        print("C4", e.run("hmm", g["obs_hmm_64"], 1 << lg)["device_ms"])
    if which in ("c5", "all"):
        print("C5 est", e.run("hmm", g["obs_hmm_1000"], 1 << (lg - 4))["device_ms"])
    if which in ("c3rows", "all"):
        print("C3 rows", e.run("linear_gaussian_1d", g["obs_linear_gaussian_32"], 1 << lg, force_rows=True)["device_ms"])
    if which in ("c5rows", "all"):
        print("C5 rows", e.run("hmm", g["obs_hmm_1000"], 1 << (lg - 4), force_rows=True)["device_ms"])
    if which in ("c4rows", "all"):
        print("C4 rows", e.run("hmm", g["obs_hmm_64"], 1 << lg, force_rows=True)["device_ms"])
