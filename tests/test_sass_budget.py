"""The headline kernel's particle loop, read from the SASS of the built library: FP64-dominated, no
local-memory traffic, and the budget file bench.py uses matches the binary."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_mix  # noqa: E402

LIB = os.path.join(ROOT, "cpprob_b200", "lib", "libcpprob_sis.so")
KERNEL = "k_sis_fusedIN6models27gaussian_unknown_mean_modelELi1"


def test_particle_loop_budget():
    b = sass_mix.loop_budget(LIB, KERNEL)
    saved = json.load(open(os.path.join(ROOT, "cpprob_b200", "lib", "sass_budget.json")))
    for k in ("total", "fp64", "dfma", "dadd", "dmul"):
        assert saved[k] == b[k], k
    # one trip = 2 particles: <= 70 FP64-pipe instructions and <= 150 instructions per particle in all
    assert b["fp64"] <= 140 and b["total"] <= 300
    # every FP64 instruction holds the issue port for two cycles ("issue model"): the static
    # ceiling of the FP64-pipe utilisation is 2F / (2F + O)
    ceiling = 2 * b["fp64"] / (2 * b["fp64"] + (b["total"] - b["fp64"]))
    assert ceiling >= 0.60


def test_no_local_memory_in_sis_kernels():
    name, body = sass_mix.kernel_sass(LIB, KERNEL)
    assert "LDL" not in body and "STL" not in body
    log = open(os.path.join(ROOT, "cpprob_b200", "lib", "models_builtin.ptxas.log")).read()
    blocks = log.split("Compiling entry function")
    fused = [b for b in blocks if "k_sis_fused" in b or "k_sis_rows" in b]
    assert fused and all("0 bytes spill stores" in b for b in fused)
