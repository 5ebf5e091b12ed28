"""The C-ABI library loads without a GPU and exports every symbol include/cpprob_sis.h declares; the
product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

from cpprob_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cpprob_sis.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cpprob_sis_[a-z_0-9]+)\s*\(", text)) - {"cpprob_sis_block_fn"})


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cpprob_sis.h but not exported"
    assert sorted(capi.SYMBOLS) == names


def test_abi_version_and_builtin_models():
    L = capi.lib()
    assert L.cpprob_sis_abi_version() == 1
    names = [L.cpprob_sis_model_name(i).decode() for i in range(L.cpprob_sis_model_count())]
    for m in ("gaussian_unknown_mean", "gaussian_unknown_mean_mu", "linear_gaussian_1d", "hmm"):
        assert m in names
    assert L.cpprob_sis_find_model(b"no_such_model") == -3
    assert b"no_such_model" in L.cpprob_sis_last_error()


def test_shard_plan_host_arithmetic():
    n = 10 * capi.CHUNK + 17                       # 11 chunks, the last one ragged
    for world in (1, 2, 3, 4, 8, 16):
        covered, chunks = 0, 0
        for r in range(world):
            cf, ncl, nct, fp, nl = capi.plan_shard(n, r, world)
            assert nct == 11 and cf == chunks and fp == cf * capi.CHUNK
            chunks += ncl
            covered += nl
        assert chunks == 11 and covered == n
    with pytest.raises(capi.SisError):
        capi.plan_shard(n, 2, 2)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SisError) as ei:
        capi.Engine()
    assert "no CPU fallback" in str(ei.value)
