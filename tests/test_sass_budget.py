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
    per = saved["particles_per_trip"]
    assert per == b["particles_per_trip"] and per in (2, 4)
    # per particle: <= 27 FP64-pipe instructions and <= 74 hot instructions in all (the ziggurat draw is one DFMA, the
    # weight's exp nine; `cold` = call set-up that only the 0.06 % slow draws execute)
    hot = b["total"] - b["cold"]
    assert b["fp64"] / per <= 27 and hot / per <= 74
    # every FP64 instruction holds the issue port for two cycles ("issue model"): issue slots per
    # particle = 2F + O
    slots = (2 * b["fp64"] + (hot - b["fp64"])) / per
    assert slots <= 100


def test_no_local_memory_in_the_particle_loop():
    import re
    name, body = sass_mix.kernel_sass(LIB, KERNEL)
    lo, hi = sass_mix.particle_loop(body)
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            assert not m.group(2).startswith(("LDL", "STL")), line
