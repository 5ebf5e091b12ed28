"""The CUDA path against committed outputs of the REFERENCE'S OWN SIS loop (tests/golden/ref_sis_golden.json, written by
oracle/_ref/ref_sis through tests/golden/make_ref_sis_golden.py): the device recomputes the log-weight of every golden trace
from its sampled values (cpprob_sis_replay) and must agree with the log-weight the reference wrote, to 1e-12 relative — with
nothing of oracle/ or /root/reference needed at test time."""
import json
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_sis_golden.json")))


@pytest.mark.parametrize("case", FX["cases"], ids=[c["model"] for c in FX["cases"]])
def test_device_log_weights_equal_the_reference_loops(engine, case):
    values = np.array([[float.fromhex(v) for v in row] for row in case["values_hex"]])
    text = case["files"].get(".real") or case["files"][".int"]
    ref_lw = np.array([float(re.search(r"\] (\S+)\)$", line).group(1)) for line in text.splitlines()])
    if ".int" in case["files"]:
        got = engine.replay(case["model"], case["obs"], int_rows=values.T.astype(np.int32))
    else:
        got = engine.replay(case["model"], case["obs"], real_rows=values.T)
    assert got.shape == ref_lw.shape
    np.testing.assert_allclose(got, ref_lw, rtol=1e-12)
