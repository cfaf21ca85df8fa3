"""The batched numpy oracle (oracle/vector_oracle.py) against the reference's
golden vectors: every lane of a golden case becomes one env of the batch, the
recorded draws are replayed, resets are issued where the reference reset."""
import warnings

import numpy as np
import pytest

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_oracle import VectorDiscreteOracle
from tests import golden_util as gu
from tests.golden.cases import CASES

CASES_D = [n for n in gu.DISCRETE_CASES
           if not CASES[n]["config"].get("image_representations")]


def replay_golden_through(vec_reset, vec_step, g):
    """Drive a batched implementation through a golden case.  `vec_reset(mask,
    reset_u)` and `vec_step(actions, transition_u, reward_noise)` are thin
    adapters; returns nothing, asserts bit-exact agreement."""
    K, T = g["done"].shape
    cur = vec_reset(None, g["init_reset_u"])
    assert np.array_equal(cur, g["init_state"])
    for t in range(T):
        obs, r, done = vec_step(g["actions"][:, t], g["transition_u"][:, t],
                                g["reward_noise"][:, t])
        assert np.array_equal(obs, g["state"][:, t]), t
        assert np.array_equal(r, g["reward"][:, t]), (t, r, g["reward"][:, t])
        assert np.array_equal(done, g["done"][:, t]), t
        m = g["reset_after"][:, t]
        if m.any():
            cur = vec_reset(m, np.nan_to_num(g["reset_u"][:, t]))
            assert np.array_equal(cur[m], g["reset_state"][m, t]), t


@pytest.mark.parametrize("name", CASES_D)
def test_vector_oracle_replays_reference_golden(name):
    g = gu.load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalar = ScalarRLToyEnv(**gu.case_config(name))
    K = g["done"].shape[0]
    vec = VectorDiscreteOracle(scalar, K)

    def vec_reset(mask, reset_u):
        return vec.reset(mask=mask, reset_u=reset_u)

    def vec_step(a, u, n):
        out = vec.rollout(1, actions=a[None], replay=dict(
            transition_u=u[None], reward_noise=n[None]))
        return out["obs"][0], out["reward"][0], out["terminated"][0]

    replay_golden_through(vec_reset, vec_step, g)


def test_philox_oracle_rollout_is_chunk_invariant():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalar = ScalarRLToyEnv(**gu.case_config("c2_every1"))
    a = VectorDiscreteOracle(scalar, 16, autoreset=True, horizon=7, seed=5)
    b = VectorDiscreteOracle(scalar, 16, autoreset=True, horizon=7, seed=5)
    a.reset(); b.reset()
    ra = a.rollout(12)
    rb1, rb2 = b.rollout(5), b.rollout(7)
    for k in ra:
        assert np.array_equal(ra[k], np.concatenate([rb1[k], rb2[k]])), k
    assert ra["terminated"].any() and ra["truncated"].any()


@pytest.mark.parametrize("p", [1e-9, 0.001, 0.1, 0.25, 1 / 3, 0.5, 0.9, 0.999, 1.0])
@pytest.mark.parametrize("S", [2, 3, 8, 13, 100, 65535])
def test_closed_form_transition_noise_is_exact_to_2_pow_minus_32(p, S):
    """The Philox-mode noisy draw (oracle/philox.py, csrc/context.cu): P(noisy)
    = T / 2^32 with T = round(p 2^32), and the S-1 other states share [0, T)
    in bins whose sizes differ from T / (S-1) by less than 2 words."""
    from oracle import philox as px
    T, M, sh, _ = px.transition_noise_params(p, S)
    assert T == int(np.floor(p * 2.0 ** 32 + 0.5)) and 0 < M < 2 ** 32 and 32 <= sh <= 63
    assert ((T - 1) * M) >> sh <= S - 2          # the index never overflows
    n = S - 1
    if T < n:
        return
    ks = np.unique(np.concatenate([np.arange(min(n, 50)), np.arange(max(n - 50, 0), n)]))
    for k in ks.tolist():
        lo = -((-k << sh) // M)                  # first w with (w M) >> sh == k
        hi = min(-((-(k + 1) << sh) // M), T)
        assert abs((hi - lo) - T / n) < 2, (k, lo, hi)
    w = np.array([0, T // 2, T - 1, T, min(T + 5, 2 ** 32 - 1), 2 ** 32 - 1], dtype=np.uint64)
    for nxt in (0, S // 2, S - 1):
        out = px.noisy_next_state(w, np.full(w.shape, nxt), (T, M, sh, S))
        assert ((out != nxt) == (w < T)).all()   # noisy <=> state changed
        assert ((0 <= out) & (out < S)).all()
