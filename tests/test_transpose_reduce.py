"""The warp epilogue of k_sis_fused (cpprob_b200/csrc/sis_kernels.cuh: warp_transpose_sum / transpose_col), emulated lane by
lane in numpy: for every column count the kernels use it must give, BIT FOR BIT, what the plain xor-butterfly gives (the tree
round 1 shipped and every committed fingerprint was computed with), and every column must end up in exactly one lane."""
import numpy as np
import pytest

LANES = np.arange(32)
OFFS = (16, 8, 4, 2, 1)


def butterfly(x):
    x = x.copy()
    for off in OFFS:
        x = x + x[LANES ^ off]
    return x


def warp_transpose_sum(v, c):
    """v[32][c] -> v[:, 0]; statement for statement the device template (P columns left, H kept per stage)."""
    v, p = v.copy(), c
    for off in OFFS:
        if p == 1:
            v[:, 0] = v[:, 0] + v[LANES ^ off, 0]
        else:
            h = (p + 1) // 2
            hi = (LANES & off) != 0
            for j in range(h):
                upper = v[:, h + j] if h + j < p else np.zeros(32)
                send = np.where(hi, v[:, j], upper)
                keep = np.where(hi, upper, v[:, j])
                v[:, j] = keep + send[LANES ^ off]
            p = h
    return v[:, 0]


def transpose_col(lane, c):
    base, valid, p, owner = 0, c, c, True
    for off in OFFS:
        hi = (lane & off) != 0
        if p == 1:
            owner = owner and not hi
        else:
            h = (p + 1) // 2
            if hi:
                base, valid = base + h, valid - h
            else:
                valid = min(valid, h)
            p = h
    return base if owner and valid >= 1 else -1


@pytest.mark.parametrize("c", [4, 6, 10, 1, 2, 3, 5, 7, 9, 16])      # 4 / 6 / 10: NR = 1 / 2 / 4 real predicts (S0, S00, S1, S2 ...)
def test_transpose_reduce_is_the_butterfly(c):
    rng = np.random.default_rng(c)
    cols = [transpose_col(int(l), c) for l in LANES]
    assert sorted(x for x in cols if x >= 0) == list(range(c))       # every column has exactly one owner lane
    for _ in range(40):
        v = rng.normal(size=(32, c)) * 10.0 ** rng.integers(-12, 12, size=(32, c))
        v[rng.random((32, c)) < 0.05] = 0.0
        got = warp_transpose_sum(v, c)
        for lane, col in enumerate(cols):
            if col >= 0:
                want = butterfly(v[:, col])
                assert (want == want[0]).all()                        # the butterfly leaves the same bits in every lane
                assert got[lane].tobytes() == want[0].tobytes()


def test_owner_lanes_of_the_shipped_instantiations():
    assert [l for l in range(32) if transpose_col(l, 4) >= 0] == [0, 8, 16, 24]
    assert [transpose_col(l, 4) for l in (0, 8, 16, 24)] == [0, 1, 2, 3]
    assert [l for l in range(32) if transpose_col(l, 6) >= 0] == [0, 4, 8, 16, 20, 24]
    assert [l for l in range(32) if transpose_col(l, 10) >= 0] == [0, 2, 4, 8, 10, 16, 18, 20, 24, 26]
