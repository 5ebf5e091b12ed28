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


def build_engine():
    _make(os.path.join(ROOT, "cpprob_b200", "csrc"))
    return os.path.join(ROOT, "cpprob_b200", "lib", "libcpprob_sis.so")


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
