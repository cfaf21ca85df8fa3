"""The batched continuous oracle against the reference's golden vectors
(recorded noise replayed, resets where the reference reset)."""
import warnings

import numpy as np
import pytest

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_continuous_oracle import VectorContinuousOracle
from tests import golden_util as gu
from tests.golden.cases import CASES

CASES_C = [n for n in gu.CONTINUOUS_CASES]


def is_exact_case(name):
    """Bit-exactness is expected except where numpy's BLAS dot over > 2
    elements enters (action_loss_weight * ||action||, 6-dim action): its
    summation order is a property of the OpenBLAS kernel of the host CPU, so
    those cases are held to the contract's 1e-5 relative tolerance instead."""
    cfg = CASES[name]["config"]
    return not (cfg.get("action_loss_weight") and cfg["state_space_dim"] > 2)


def replay_continuous_golden(vec_reset, vec_step, g, exact=True, rtol=1e-5):
    """Drive a batched implementation through a continuous golden case."""
    K, T = g["done"].shape
    cur = vec_reset(None, g["init_state"])
    assert np.array_equal(cur, g["init_state"])
    worst = 0.0
    for t in range(T):
        sn = g["state_noise"][:, t] if "state_noise" in g else None
        obs, r, done, derivs = vec_step(
            g["actions"][:, t], None if sn is None else np.nan_to_num(sn),
            np.nan_to_num(g["reward_noise"][:, t]))
        want_r = g["reward"][:, t].astype(np.float32)
        if exact:
            assert np.array_equal(obs, g["state"][:, t]), t
            assert np.array_equal(r, want_r), (t, r, want_r)
            if derivs is not None:
                assert np.array_equal(derivs, g["derivs"][:, t].astype(derivs.dtype)), t
        else:
            np.testing.assert_allclose(obs, g["state"][:, t], rtol=rtol, atol=1e-7)
            np.testing.assert_allclose(r, want_r, rtol=rtol, atol=1e-6)
        assert np.array_equal(done, g["done"][:, t]), t
        m = g["reset_after"][:, t]
        if m.any():
            cur = vec_reset(m, g["reset_state"][:, t])
            assert np.array_equal(cur[m], g["reset_state"][m, t]), t
    return worst


@pytest.mark.parametrize("name", CASES_C)
def test_vector_continuous_oracle_replays_reference_golden(name):
    g = gu.load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalar = ScalarRLToyEnv(**gu.case_config(name))
    K = g["done"].shape[0]
    vec = VectorContinuousOracle(scalar, K)

    def vec_reset(mask, init):
        return vec.reset(mask=mask, init_state=init)

    def vec_step(a, sn, rn):
        out = vec.rollout(1, actions=a[None], replay=dict(
            state_noise=None if sn is None else sn[None], reward_noise=rn[None]))
        return (out["obs"][0], out["reward"][0], out["terminated"][0],
                np.transpose(vec.sd, (1, 0, 2)))

    replay_continuous_golden(vec_reset, vec_step, g, exact=is_exact_case(name))
