#!/usr/bin/env python3
"""A small pass over every path for compute-sanitizer (memcheck / racecheck / initcheck): fused, staged (reals, ints, packed
ints, accumulators in L2), rows with both reductions, emission with the text stage, replay, reduce_records, and (on >= 2
GPUs) the peer-memory exchange of run_multi.
usage: compute-sanitizer --tool racecheck python tools/sanitize_once.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import analytic  # noqa: E402
from cpprob_b200 import Engine, capi  # noqa: E402

g = analytic.golden()
n = 4096 + 700
with Engine(seed=7, max_batch=capi.CHUNK) as e, tempfile.TemporaryDirectory() as tmp:
    for model, obs in (("gaussian_unknown_mean", [3.0, 4.0]), ("linear_gaussian_1d", g["obs_linear_gaussian_32"]),
                       ("linear_gaussian_1d", (g["obs_linear_gaussian_32"] * 2)[:45]), ("hmm", g["obs_hmm_64"]),
                       ("hmm", g["obs_hmm_1000"][:300]), ("hmm", g["obs_hmm_1000"]), ("all_distr", [0.0, 0.0])):
        a = e.run(model, obs, n)
        b = e.run(model, obs, n, force_rows=True)
        print(model, len(obs), a["path"], b["path"], bool((a["sums"] == b["sums"]).all()) if a["path"] != "fused" else "-")
        e.infer_to_files(model, obs, 1500, os.path.join(tmp, model + str(len(obs))))
    rec = e.run("linear_gaussian_1d", g["obs_linear_gaussian_32"][:5], 3000, collect=True)
    e.replay("linear_gaussian_1d", g["obs_linear_gaussian_32"][:5], real_rows=rec["real_rows"])
    e.reduce_records(rec["log_w"], real_rows=rec["real_rows"])
    rec = e.run("hmm", g["obs_hmm_64"][:9], 3000, collect=True)
    e.reduce_records(rec["log_w"], int_rows=rec["int_rows"])
    e.run_dist("hmm", g["obs_hmm_64"][:9], 3000)
# the multi-GPU exchange over peer memory (k_fold_units / k_push_rows into the peers' windows, the merge
# kernel's wait on the epoch flags) on the fused, staged and row-fed shapes, three inferences each (both gather buffers)
import torch  # noqa: E402
if True:
    # two GPUs if there are two, else two ranks on one GPU (the exchange kernels are the same)
    engines = [Engine(device=d if torch.cuda.device_count() >= 2 else 0, seed=7) for d in range(2)]
    try:
        for model, obs, m in (("gaussian_unknown_mean", [3.0, 4.0], 5 * capi.CHUNK + 77), ("hmm", g["obs_hmm_64"][:9], 3 * capi.CHUNK + 5),
                              ("linear_gaussian_1d", g["obs_linear_gaussian_32"][:6], 3 * capi.CHUNK + 5)):
            for _ in range(3):
                st = capi.run_multi(engines, model, obs, m)
            print("run_multi", model, engines[0].comm_exchange(), st["path"])
    finally:
        for x in engines:
            x.close()
print("sanitize pass done")
