"""Builds the in-tree native artefacts: libcpprob_sis.so (nvcc, sm_100a) and the oracle (g++)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make(directory, *targets):
    env = dict(os.environ)
    env.setdefault("PATH", "")
    if "/usr/local/cuda/bin" not in env["PATH"]:
        env["PATH"] = "/usr/local/cuda/bin:" + env["PATH"]
    subprocess.run(["make", "-j4", *targets], cwd=directory, check=True, env=env)


HEADLINE_KERNEL = "k_sis_fusedIN6models27gaussian_unknown_mean_modelELi1"


def write_sass_budget(lib):
    """Static instruction budget of the headline kernel's particle loop (one trip = 2 particles), read from
    the SASS of the library just built; bench.py turns it into the roofline's flop / instruction counts."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_mix
    b = sass_mix.loop_budget(lib, HEADLINE_KERNEL)
    b["kernel"] = HEADLINE_KERNEL
    with open(os.path.join(os.path.dirname(lib), "sass_budget.json"), "w") as f:
        json.dump(b, f)
    return b


def build_engine():
    _make(os.path.join(ROOT, "cpprob_b200", "csrc"))
    lib = os.path.join(ROOT, "cpprob_b200", "lib", "libcpprob_sis.so")
    write_sass_budget(lib)
    return lib


def build_oracle():
    _make(os.path.join(ROOT, "oracle"))
    return os.path.join(ROOT, "oracle", "liboracle.so")


def build_examples():
    ex = os.path.join(ROOT, "examples")
    if os.path.exists(os.path.join(ex, "Makefile")):
        _make(ex)


if __name__ == "__main__":
    print(build_engine())
    print(build_oracle())
    build_examples()
