"""C++14 host API (include/cpprob): builds warning-free with a plain C++14 compiler, and its GPU-free
parts (Philox host twin, structure probe, serialization grammar, error behaviour) hold."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


@pytest.fixture(scope="module")
def built():
    r = subprocess.run(["make", "-C", EX], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "warning" not in (r.stdout + r.stderr)
    return os.path.join(EX, "bin")


def test_host_twin_check(built):
    r = subprocess.run([os.path.join(built, "host_twin_check")], capture_output=True, text=True)
    assert r.returncode == 0 and "host twin check: ok" in r.stdout, r.stdout + r.stderr


def test_limits_fail_loudly(built):
    """discrete_distribution with more weights than its inline capacity throws instead of sampling another distribution."""
    r = subprocess.run([os.path.join(built, "limits_check")], capture_output=True, text=True)
    assert r.returncode == 0 and "more weights than its capacity" in r.stdout, r.stdout + r.stderr


def test_cli_argument_errors(built):
    main = os.path.join(built, "main")
    r = subprocess.run([main, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--n_samples" in r.stdout
    r = subprocess.run([main, "--sis"], capture_output=True, text=True)
    assert r.returncode != 0 and "'--model' is required" in r.stderr
    r = subprocess.run([main, "--sis", "--model", "nope"], capture_output=True, text=True)
    assert r.returncode != 0 and "Model not available." in r.stderr
    r = subprocess.run([main, "--csis", "--model", "hmm"], capture_output=True, text=True)
    assert r.returncode != 0 and "inference compilation" in r.stderr


@pytest.mark.gpu
def test_readme_program_on_gpu(built, tmp_path):
    """README.md:102-116 compiled unchanged against the C++14 API, run on the GPU."""
    env = dict(os.environ, CPPROB_SIS_SEED="12345")
    r = subprocess.run([os.path.join(built, "hello_sis")], capture_output=True, text=True, cwd=tmp_path, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "Estimators for posterior_sis.real" and lines[1] == "Mean:"
    mean = float(lines[2].split("Mean: ")[1])
    var = float(lines[3].split("Variance: ")[1])
    assert abs(mean - 2.32353) < 4 * 0.013 and abs(var - 1.05882) < 0.1
    assert sorted(os.listdir(tmp_path)) == ["posterior_sis.ids", "posterior_sis.real", "posterior_sis.stats"]
    assert len(open(tmp_path / "posterior_sis.real").read().splitlines()) == 10_000
    # the same seed reproduces the run bit for bit; StatsPrinter of the oracle prints the same text
    r2 = subprocess.run([os.path.join(built, "hello_sis")], capture_output=True, text=True, cwd=tmp_path, env=env)
    assert r2.stdout == r.stdout or len(open(tmp_path / "posterior_sis.real").read().splitlines()) == 20_000


@pytest.mark.gpu
def test_stats_printer_text_equals_reference(built, tmp_path, oracle):
    env = dict(os.environ, CPPROB_SIS_SEED="7")
    main = os.path.join(built, "main")
    obs = "[" + " ".join(f"{0.1 * i:.3f}" for i in range(10)) + "]"
    r = subprocess.run([main, "--sis", "--estimate", "--model", "hmm", "-n", "5000", "-o", obs, "--model_folder", str(tmp_path / "hmm")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("Sequential Importance Sampling (SIS)\nPosterior Distribution Estimators\n")
    ours = r.stdout.split("Posterior Distribution Estimators\n", 1)[1]
    ref = oracle.stats_text(str(tmp_path / "hmm" / "post_sis"))
    assert ours.rstrip("\n") == ref.rstrip("\n")
    r = subprocess.run([main, "--sis", "--estimate", "--model", "unk_mean", "-n", "5000", "-o", "8 9", "--model_folder", str(tmp_path / "um")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    ours = r.stdout.split("Posterior Distribution Estimators\n", 1)[1]
    ref = oracle.stats_text(str(tmp_path / "um" / "post_sis"))
    assert ours.splitlines()[:2] == ref.splitlines()[:2]          # header + "Mu:"
    for a, b in zip(ours.splitlines()[2:4], ref.splitlines()[2:4]):
        assert a.split(": ")[0] == b.split(": ")[0] and abs(float(a.split(": ")[1]) - float(b.split(": ")[1])) < 1e-4
    # estimators-only mode: no record files, StatsPrinter falls back to the .stats sidecar
    env2 = dict(env, CPPROB_SIS_EMIT="none")
    r = subprocess.run([main, "--sis", "--estimate", "--model", "linear_gaussian", "-n", "200000", "-o",
                        "[" + " ".join(["0.5"] * 50) + "]", "--model_folder", str(tmp_path / "lg")], capture_output=True, text=True, env=env2)
    assert r.returncode == 0, r.stdout + r.stderr
    assert sorted(os.listdir(tmp_path / "lg")) == ["post_sis.ids", "post_sis.stats"]
    assert "State 49:\n  Mean: " in r.stdout


@pytest.mark.gpu
def test_cli_models_with_aggregate_observations(built, tmp_path):
    """Observation strings of the aggregate-typed models go through the serialization grammar
    (poly_adjustment.hpp:33: -o [[1 2.1] [2 3.9] ...])."""
    env = dict(os.environ, CPPROB_SIS_SEED="3")
    main = os.path.join(built, "main")
    pts = "[[1 2.1] [2 3.9] [3 5.3] [4 7.7] [5 10.2] [6 12.9]]"
    r = subprocess.run([main, "--sis", "--estimate", "--model", "linear_regression", "-n", "20000", "-o", pts,
                        "--model_folder", str(tmp_path / "lr")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Coefficient 0:\n  Mean: " in r.stdout and "Coefficient 1:\n  Mean: " in r.stdout
    r = subprocess.run([main, "--sis", "--estimate", "--model", "dyn_linear_reg", "-n", "20000", "-o",
                        "[(1 2.1) (2 3.9) (3 5.3)]", "--model_folder", str(tmp_path / "dl")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "a:\n  Mean: " in r.stdout and "b:\n  Mean: " in r.stdout
    r = subprocess.run([main, "--sis", "--estimate", "--model", "unk_mean_2d", "-n", "20000", "-o", "[3 4]",
                        "--model_folder", str(tmp_path / "g2")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Mu:\n  Mean: [" in r.stdout and "  Variance: [" in r.stdout
    r = subprocess.run([main, "--sis", "--estimate", "--model", "unk_mean_rejection", "-n", "20000", "-o", "3 4",
                        "--model_folder", str(tmp_path / "rj")], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "Mu:\n  Mean: 3." in r.stdout, r.stdout + r.stderr
    r = subprocess.run([main, "--sis", "--model", "unk_mean", "-o", "3 oops", "--model_folder", str(tmp_path / "bad")],
                       capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "Could not parse the observations." in r.stderr


@pytest.mark.gpu
def test_repeated_inference_reuses_the_engine(built, tmp_path):
    """cpprob::inference called again and again from one process keeps its engine (sis::cached_engine): after the first
    call, a 10,000-particle README inference with its posterior file costs well under 5 ms (the kernels take ~30 us; the
    rest is one stream synchronisation per stage and the file append)."""
    import re
    prog = os.path.join(built, "repeat_inference")
    r = subprocess.run([prog, str(tmp_path / "p"), "12", "10000"], capture_output=True, text=True, env=dict(os.environ, CPPROB_SIS_SEED="9"))
    assert r.returncode == 0, r.stdout + r.stderr
    ms = [float(x) for x in re.findall(r"call \d+: ([0-9.]+) ms", r.stdout)]
    assert len(ms) == 12
    later = sorted(ms[2:])
    assert later[len(later) // 2] < 5.0, ms
    assert ms[0] > 5 * later[len(later) // 2], ms            # the first call paid for the context and the tables
    assert "Mean:\n  Mean: 2.3" in r.stdout
    assert len(open(tmp_path / "p.real").read().splitlines()) == 10000
