#!/usr/bin/env python3
"""Statistical calibration of the sampler + estimator: z-scores of the posterior-mean / log-evidence estimates of
the README model over independent seeds must look like N(0,1).  usage: python tools/calibration.py [n_seeds] [n]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpprob_b200 import Engine  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 250_000_000
zs, zl = [], []
for s in range(n_seeds):
    with Engine(seed=1000 + s) as e:
        st = e.run("gaussian_unknown_mean", [3.0, 4.0], n)
    zs.append((st["real_mean"][0] - 2.323529411764706) / (1.2973 / math.sqrt(n)))
    zl.append((st["log_evidence"] + 4.398851446364485) / (0.9803 / math.sqrt(n)))   # sd of w/E[w] = sqrt(1/0.50992 - 1)
m = sum(zs) / len(zs)
v = sum((z - m) ** 2 for z in zs) / (len(zs) - 1)
print("mean z:", " ".join(f"{z:+.2f}" for z in zs))
print(f"mean-of-z {m:+.3f} (expect 0 +- {1 / math.sqrt(n_seeds):.3f}), var-of-z {v:.3f} (expect 1)")
ml = sum(zl) / len(zl)
print("logZ z:", " ".join(f"{z:+.2f}" for z in zl))
print(f"mean-of-z {ml:+.3f}, var-of-z {sum((z - ml) ** 2 for z in zl) / (len(zl) - 1):.3f}")
