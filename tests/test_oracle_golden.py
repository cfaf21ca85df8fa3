"""The scalar oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Two legs per case:

  numpy leg   the oracle consumes its own PCG64 streams, seeded like the
              reference's -> every state / reward / image equal bit for bit;
  replay leg  the oracle is fed the *recorded* draws instead.
"""
import warnings

import numpy as np
import pytest

from oracle.scalar_env import ReplayDraws, ScalarRLToyEnv, np_random
from tests import golden_util as gu
from tests.golden.cases import CASES


def _tables_equal(env, g, cfg):
    if cfg["state_space_type"] != "discrete":
        return
    assert np.array_equal(np.array(env.transition_matrix, dtype=np.int64), g["P"])
    assert np.array_equal(np.array(env.terminal_states), g["terminal_states"])
    assert np.array_equal(env.init_state_dist, g["init_state_dist"])
    if cfg.get("irrelevant_features"):
        assert np.array_equal(np.array(env.transition_matrix_irr, dtype=np.int64),
                              g["P_irr"])
    if not cfg.get("use_custom_mdp"):
        gold = gu.golden_sequences(g)
        assert list(env.rewardable_sequences.items()) == list(gold.items())
    assert int(env.reward_every_n_steps) == int(g["reward_every_n_steps"])


def _check_lane(env, g, k, cfg, H):
    cont = cfg["state_space_type"] == "continuous"
    image = bool(cfg.get("image_representations"))
    T = g["done"].shape[1]
    for t in range(T):
        a = g["actions"][k, t]
        a = a.copy() if (cont or np.ndim(a)) else int(a)
        obs, r, done, trunc, _ = env.step(a)
        assert np.array_equal(env.curr_state, g["state"][k, t]), (k, t)
        assert float(r) == g["reward"][k, t], (k, t, r, g["reward"][k, t])
        assert isinstance(r, np.float32) == bool(g["reward_is_f32"][k, t])
        assert done == bool(g["done"][k, t])
        if cont:
            assert np.array_equal(np.array(env.state_derivatives),
                                  g["derivs"][k, t])
        if image:
            assert np.array_equal(obs, g["obs_image"][k, t]), (k, t)
            if not cont:
                p = env.last_image_params  # of the last sub-image drawn
                got = [p["R"], p["shift_w"], p["shift_h"],
                       -1 if p["rotation"] is None else p["rotation"], p["flip"]]
                assert got == list(g["image_params"][k, t].reshape(-1, 5)[-1])
        do_reset = done or t % H == H - 1
        assert do_reset == bool(g["reset_after"][k, t])
        if do_reset:
            obs_r, _ = env.reset()
            assert np.array_equal(env.curr_state, g["reset_state"][k, t])
            if image:
                assert np.array_equal(obs_r, g["reset_image"][k, t])


NOT_GRID = [n for n in CASES if n not in gu.GRID_CASES]  # grid: tests/test_grid.py


def numpy_leg(g, cfg, H):
    """The oracle on its own PCG64 streams against one golden record `g`
    (also used on freshly generated records by tests/test_fuzz_reference.py)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = ScalarRLToyEnv(**cfg)
    _tables_equal(env, g, cfg)
    cont = cfg["state_space_type"] == "continuous"
    for k in range(g["done"].shape[0]):
        s = gu.lane_seed(k)
        if not cont:
            env.rng_S, _ = np_random(s + 1)
            if cfg.get("irrelevant_features"):
                env.rng_S1, _ = np_random(s + 4)
        else:
            env.rng_F, _ = np_random(s + 2)
        if cfg.get("image_representations") and not cont:
            env.rng_I, _ = np_random(s + 3)
        obs0, _ = env.reset(seed=s)
        assert np.array_equal(env.curr_state, g["init_state"][k])
        if cfg.get("image_representations"):
            assert np.array_equal(obs0, g["init_image"][k])
        _check_lane(env, g, k, cfg, H)


@pytest.mark.parametrize("name", NOT_GRID)
def test_oracle_numpy_streams_match_reference_golden(name):
    numpy_leg(gu.load(name), gu.case_config(name),
              CASES[name].get("horizon", 12))


def _lane_feed(g, k, cont, image_params=None):
    """Recorded draws of lane k in consumption order."""
    T = g["done"].shape[1]
    feed = {"transition_u": [], "reward_noise": [], "reset_u": [],
            "state_noise": [], "reset_state": [], "image_scale_u": [],
            "image_int": [], "irr_transition_u": [], "irr_reset_u": []}
    irr = "irr_reset_u" in g
    if cont:
        feed["reset_state"].append(g["init_state"][k])
    else:
        feed["reset_u"].append(float(g["init_reset_u"][k]))
        if irr:
            feed["irr_reset_u"].append(float(g["init_irr_reset_u"][k]))
    for t in range(T):
        if not np.isnan(g["transition_u"][k, t]):
            feed["transition_u"].append(float(g["transition_u"][k, t]))
        if not np.isnan(g["reward_noise"][k, t]):
            feed["reward_noise"].append(float(g["reward_noise"][k, t]))
        if irr and not np.isnan(g["irr_transition_u"][k, t]):
            feed["irr_transition_u"].append(float(g["irr_transition_u"][k, t]))
        if cont and not np.isnan(g["state_noise"][k, t]).any():
            feed["state_noise"].append(g["state_noise"][k, t])
        if g["reset_after"][k, t]:
            if cont:
                feed["reset_state"].append(g["reset_state"][k, t])
            else:
                feed["reset_u"].append(float(g["reset_u"][k, t]))
                if irr:
                    feed["irr_reset_u"].append(float(g["irr_reset_u"][k, t]))
    return feed


@pytest.mark.parametrize("name", [n for n in NOT_GRID
                                  if not CASES[n]["config"].get(
                                      "image_representations")])
def test_oracle_replay_of_recorded_draws(name):
    replay_leg(gu.load(name), lambda: gu.case_config(name),
               CASES[name].get("horizon", 12))


def replay_leg(g, make_cfg, H):
    """The oracle fed with the draws recorded in `g`; `make_cfg()` returns a
    fresh config dict (the constructor edits it in place)."""
    cfg = make_cfg()
    cont = cfg["state_space_type"] == "continuous"
    for k in range(g["done"].shape[0]):
        feed = _lane_feed(g, k, cont)
        # the ctor itself performs one reset: give it a throw-away draw
        if cont:
            feed["reset_state"].insert(0, g["init_state"][k])
        else:
            feed["reset_u"].insert(0, 0.0)
            if cfg.get("irrelevant_features"):
                feed["irr_reset_u"].insert(0, 0.0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = ScalarRLToyEnv(draws=ReplayDraws(feed), **make_cfg())
        env.reset()
        assert np.array_equal(env.curr_state, g["init_state"][k])
        _check_lane(env, g, k, cfg, H)
        assert all(len(v) == 0 for v in env.draws.feed.values())
