"""CPU check of the device text stage's `%.15e` formatter (cpprob_b200/csrc/text_format.cuh is
__host__ __device__): byte-for-byte against glibc printf — which is what the reference's
`os << std::scientific << std::setprecision(15)` (src/cpprob/state.cpp:262-267) prints — over random bit
patterns, the magnitudes the engine prints, integers, exact ties at the 16th digit, powers of two and ten,
subnormals and the specials.  The same source is compiled for the GPU; tests/test_files_gpu.py checks the
GPU-written files against the host writer."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_format_e15_matches_printf():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples"), "bin/text_format_check"], check=True)
    out = subprocess.run([os.path.join(ROOT, "examples", "bin", "text_format_check"), "400000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"checked (\d+) values: (\d+) mismatches, (\d+) ambiguous", out.stdout)
    assert m and int(m.group(1)) > 2_000_000 and int(m.group(2)) == 0
    # the undecidable window has relative width 2^-74: never hit by these inputs
    assert int(m.group(3)) == 0


def test_pow10_table_is_reproducible(tmp_path):
    """pow10_table.inc is exactly what tools/gen_pow10.py generates (no hand edits)."""
    path = os.path.join(ROOT, "cpprob_b200", "csrc", "pow10_table.inc")
    out = str(tmp_path / "pow10_table.inc")
    subprocess.run(["python", os.path.join(ROOT, "tools", "gen_pow10.py"), out], check=True, capture_output=True)
    assert open(out).read() == open(path).read()
