#!/usr/bin/env python3
"""Generates tests/golden/ref_sis_golden.json: posterior files written by the REFERENCE'S OWN SIS loop
(oracle/_ref/ref_sis = cpprob::inference(StateType::sis, ...) of /root/reference, compiled unmodified by oracle/Makefile) for a
few small cases, on prescribed sampled values (seeded here).  The fixture lets the oracle-vs-reference check run where
oracle/_ref was never built (no /root/reference).  Run in the build container:  python tests/golden/make_ref_sis_golden.py"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import analytic  # noqa: E402
import ref_lib  # noqa: E402

G = analytic.golden()
PTS = [1, 2.1, 2, 3.9, 3, 5.3, 4, 7.7, 5, 10.2, 6, 12.9]
CASES = [("gaussian_unknown_mean", [3.0, 4.0], 1, "real", 40), ("gaussian_unknown_mean_mu", [3.0, 4.0], 1, "real", 20),
         ("linear_gaussian_1d", G["obs_linear_gaussian_32"][:8], 8, "real", 20), ("hmm", G["obs_hmm_64"][:12], 12, "state", 20),
         ("gaussian_2d_unk_mean", [1.5, 2.5], 2, "real", 20), ("poly_adjustment_2", PTS, 3, "real", 20)]

out = {"generator": "tests/golden/make_ref_sis_golden.py", "what": ref_lib.load().describe(), "cases": []}
assert ref_lib.available() and os.path.exists(ref_lib.REF_SIS), "build oracle/_ref first (needs /root/reference)"
for i, (model, obs, per, kind, n) in enumerate(CASES):
    rng = np.random.default_rng(1000 + i)
    values = rng.integers(0, 3, (n, per)).astype(np.float64) if kind == "state" else rng.normal(0.5, 2.0, (n, per))
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "p")
        ref_lib.ref_sis(model, obs, n, prefix, replay=values)
        files = {ext: open(prefix + ext).read() for ext in (".real", ".int", ".any", ".ids") if os.path.exists(prefix + ext)}
    out["cases"].append({"model": model, "obs": [float(x) for x in obs], "values_hex": [[float(v).hex() for v in row] for row in values], "files": files})
with open(os.path.join(HERE, "ref_sis_golden.json"), "w") as f:
    json.dump(out, f, indent=0)
print("wrote", os.path.join(HERE, "ref_sis_golden.json"), os.path.getsize(os.path.join(HERE, "ref_sis_golden.json")), "bytes")
