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
