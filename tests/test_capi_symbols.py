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
    assert L.cpprob_sis_abi_version() == 3
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


def test_super_chunk_rows_host_arithmetic():
    """Large runs hand on super-chunk rows: 2^k chunks each, k from the run size only, ranks own whole super-chunks."""
    n = 100_000 * capi.CHUNK + 123                 # 100001 chunks -> S = 32, 3126 rows
    rows = {}
    for world in (1, 2, 3, 8):
        first_expected, covered = 0, 0
        for r in range(world):
            rf, nl, nt = capi.plan_rows(n, r, world)
            cf, ncl, nct, fp, nloc = capi.plan_shard(n, r, world)
            assert nt == 3126 and nct == 100001 and rf == first_expected
            assert cf == rf * 32 and fp == cf * capi.CHUNK          # shard boundaries sit on super-chunk boundaries
            assert ncl == min(nct, (rf + nl) * 32) - cf
            first_expected += nl
            covered += nloc
        assert first_expected == 3126 and covered == n
        rows[world] = first_expected
    # small runs: one row per chunk (or per chunk for the 8-rows-per-chunk row path, folded on the rank)
    assert capi.plan_rows(10 * capi.CHUNK + 17, 0, 1) == (0, 11, 11)
    assert capi.plan_rows(10 * capi.CHUNK + 17, 1, 2, 8) == (5, 6, 11)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SisError) as ei:
        capi.Engine()
    assert "no CPU fallback" in str(ei.value)
