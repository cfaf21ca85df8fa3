"""Grid environments (state_space_type="grid"; rl_toy_env.py:1727-1778
transition, :1947-1965 Manhattan reward, :2325-2345 reset; SURVEY.md 8f row
N1).  CPU: the scalar oracle against goldens recorded from the unmodified
reference (own PCG64 streams and replayed draws).  GPU: the CUDA path against
those goldens and against the oracle.  Everything is integer / fp64 with the
reference's operation order: bit-exact, except Philox-mode rewards (1e-12).

Reference quirks reproduced on purpose (each probed on the live reference):
terminal cells never terminate (float array tested against an int64 Box);
reset() can put a coordinate ONE PAST the grid (int Box samples [0, shape]);
delay / sequence_length other than 0 / 1 crash in the reference -> rejected."""
import warnings

import numpy as np
import pytest

from oracle.scalar_env import ReplayDraws, ScalarRLToyEnv, np_random
from tests import golden_util as gu
from tests.golden.cases import CASES


def scalar_oracle(cfg, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ScalarRLToyEnv(**kw, **cfg)


def _check_lane(env, g, k, H, image):
    T = g["done"].shape[1]
    for t in range(T):
        obs, r, done, trunc, _ = env.step([int(a) for a in g["actions"][k, t]])
        assert np.array_equal(env.curr_state, g["state"][k, t]), (k, t)
        assert float(r) == g["reward"][k, t], (k, t)
        assert done == bool(g["done"][k, t])
        if image:
            assert np.array_equal(obs, g["obs_image"][k, t]), (k, t)
        do_reset = (done and k % 2 == 0) or t % H == H - 1
        assert do_reset == bool(g["reset_after"][k, t])
        if do_reset:
            obs_r, _ = env.reset()
            assert np.array_equal(env.curr_state, g["reset_state"][k, t])
            if image:
                assert np.array_equal(obs_r, g["reset_image"][k, t])


@pytest.mark.parametrize("name", gu.GRID_CASES)
def test_oracle_numpy_streams_match_reference_golden(name):
    g, cfg = gu.load(name), gu.case_config(name)
    env = scalar_oracle(cfg)
    image = bool(cfg.get("image_representations"))
    for k in range(g["done"].shape[0]):
        s = gu.lane_seed(k)
        env.rng_F, _ = np_random(s + 2)
        env.rng_A, _ = np_random(s + 5)
        obs0, _ = env.reset(seed=s)
        assert np.array_equal(env.curr_state, g["init_state"][k])
        if image:
            assert np.array_equal(obs0, g["init_image"][k])
        _check_lane(env, g, k, CASES[name].get("horizon", 12), image)


@pytest.mark.parametrize("name", [n for n in gu.GRID_CASES
                                  if not CASES[n]["config"].get("image_representations")])
def test_oracle_replay_of_recorded_draws(name):
    g, cfg = gu.load(name), gu.case_config(name)
    K, T = g["done"].shape
    for k in range(K):
        att = g["grid_attempts"]
        feed = {"reset_state": [g["init_state"][k], g["init_state"][k]]
                + [g["reset_state"][k, t] for t in range(T) if g["reset_after"][k, t]],
                "grid_noise_u": [float(u) for u in g["grid_noise_u"][k] if not np.isnan(u)],
                "grid_noise_action": [(int(i), int(v)) for kk, t, i, v in att if kk == k],
                "reward_noise": [float(n) for n in g["reward_noise"][k] if not np.isnan(n)]}
        env = scalar_oracle(cfg, draws=ReplayDraws(feed))
        env.reset()
        _check_lane(env, g, k, CASES[name].get("horizon", 12), False)
        assert all(len(v) == 0 for v in env.draws.feed.values())
