"""The fp64 ziggurat behind `normal_precision="fp64"` (CPU side, no GPU).

Pins, in order: (1) the oracle's restatement of numpy's
`random_standard_normal` and the shipped tables against numpy itself, bit for
bit, by feeding it PCG64's own raw words; (2) the tables compiled into
libmdpp_b200.so against the oracle's copy; (3) the table construction redone
with 60-digit arithmetic (numpy's carry a few ulps of generation error);
(4) the counter-based word supply of oracle/philox.ziggurat_normal."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox as px
from oracle import ziggurat as zg


@pytest.mark.parametrize("seed", [0, 12345])
def test_restatement_reproduces_numpy_standard_normal_bit_for_bit(seed):
    n = 400_000  # ~6000 wedge tests, ~100 tail draws
    want = np.random.Generator(np.random.PCG64(seed)).standard_normal(n)
    raw = iter(np.random.PCG64(seed).random_raw(2 * n).tolist())
    got = np.array([zg.standard_normal_stream(lambda: next(raw)) for _ in range(n)])
    assert np.array_equal(got, want)
    assert (np.abs(want) > zg.ZIG_R).sum() > 50  # the tail branch was exercised


def test_reference_reward_noise_call_is_this_algorithm():
    """rl_toy_env.py:1982 draws `np_random.normal(0, sigma)`: numpy computes
    loc + scale * standard_normal() -- the form the kernels use (sigma * z)."""
    sigma = 0.25
    a = np.random.Generator(np.random.PCG64(3)).normal(0, sigma, size=1000)
    b = 0 + sigma * np.random.Generator(np.random.PCG64(3)).standard_normal(1000)
    assert np.array_equal(a, b)


def test_library_tables_equal_the_oracle_copy():
    from mdp_playground_b200 import _lib
    lib = _lib.load()
    ptrs = [C.c_void_p() for _ in range(3)]
    lib.mdpp_ziggurat_tables(*[C.byref(p) for p in ptrs])
    arrs = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(256,))
            for p in ptrs]
    ki, wi, fi = zg.tables()
    assert np.array_equal(arrs[0], ki)
    assert np.array_equal(arrs[1], wi.view(np.uint64))
    assert np.array_equal(arrs[2], fi.view(np.uint64))


def test_tables_agree_with_high_precision_construction():
    ki, wi, fi = zg.tables()
    k2, w2, f2 = zg.make_tables_highprec()
    assert np.abs(wi / w2 - 1).max() < 1e-13
    assert np.abs(fi / f2 - 1).max() < 1e-14
    assert np.abs(ki.astype(np.int64) - k2.astype(np.int64)).max() < 64  # of 2^52
    x, ok = zg.fast_path(np.random.PCG64(1).random_raw(200_000))
    assert 0.984 < ok.mean() < 0.987  # first-attempt acceptance


def test_counter_based_supply_is_numpys_algorithm_on_philox_words():
    """ziggurat_normal(seed, env, step) == standard_normal_stream fed with the
    documented word sequence; even/odd steps share one Philox call."""
    seed, gids = 99, np.arange(3000, dtype=np.uint32) + 17
    for step in (0, 1, 6, 2**33 + 5):
        got = px.ziggurat_normal(seed, gids, step)
        pair = step >> 1
        w = px.philox4x32_10(gids, pair & 0xFFFFFFFF, pair >> 32, px.STREAM_ZIG, seed)
        lo, hi = (w[2], w[3]) if step & 1 else (w[0], w[1])
        for i in (0, 1, 2, 77, 2999):
            words = [(int(hi[i]) << 32) | int(lo[i])]
            for c in range(4):
                q = px.philox4x32_10(gids[i:i + 1], step & 0xFFFFFFFF, step >> 32,
                                     px.STREAM_ZIG_RETRY + c, seed)
                words += [(int(q[1][0]) << 32) | int(q[0][0]),
                          (int(q[3][0]) << 32) | int(q[2][0])]
            it = iter(words)
            assert got[i] == zg.standard_normal_stream(lambda: next(it))
    z = np.concatenate([px.ziggurat_normal(5, gids, s) for s in range(40)])
    from scipy import stats
    assert stats.kstest(z, "norm").pvalue > 1e-3   # 120 000 draws vs N(0,1)
