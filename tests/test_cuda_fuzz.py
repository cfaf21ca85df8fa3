"""Randomised configurations on the GPU: the CUDA kernels (through the C ABI)
against the batched CPU oracles on the configurations of tests/fuzz_configs.py
-- the same ones tests/test_fuzz_reference.py pins the oracle on against the
unmodified reference.  Native Philox noise on both sides, device-drawn and
given actions, ragged batch sizes, both the run-time specialised (NVRTC) and
the ahead-of-time kernels, the standard bench signature (out=, no final_obs)
as well as the generic one.

Bars: discrete / grid -- states, flags bit-exact, fp64 rewards 1e-12;
continuous -- states 1e-5 relative (north_star), rewards 1e-4, flags equal.
"""
import copy
import warnings

import numpy as np
import pytest
import torch

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_continuous_oracle import VectorContinuousOracle
from oracle.vector_grid_oracle import VectorGridOracle
from oracle.vector_oracle import VectorDiscreteOracle
from tests.fuzz_configs import (CONTINUOUS_SEEDS, DISCRETE_SEEDS, GRID_SEEDS,
                                IMAGE_SEEDS, continuous_fuzz_config,
                                discrete_fuzz_config, grid_fuzz_config,
                                image_fuzz_config)
from tests.golden.cases import materialise

pytestmark = pytest.mark.gpu


def make_env(*a, **k):
    from mdp_playground_b200 import VectorRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorRLToyEnv(*a, **k)


def scalar_oracle(cfg):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ScalarRLToyEnv(**cfg)


def _shape_of(seed):
    """Batch size (ragged on purpose), rollout length, horizon, JIT on/off."""
    r = np.random.default_rng(77 + seed)
    N = int(r.choice([1, 31, 65, 257, 700, 1500]))
    T = int(r.choice([9, 24, 41]))
    horizon = int(r.choice([0, 5, 13]))
    return N, T, horizon, bool(seed % 2)


@pytest.mark.parametrize("seed", DISCRETE_SEEDS)
def test_discrete_fuzz_cuda_vs_oracle(seed):
    cfg = discrete_fuzz_config(seed)
    N, T, horizon, jit = _shape_of(seed)
    irr = bool(cfg.get("irrelevant_features"))
    ora = VectorDiscreteOracle(scalar_oracle(materialise(copy.deepcopy(cfg))), N,
                               autoreset=True, horizon=horizon, seed=seed,
                               env_id_offset=12345)
    env = make_env(N, autoreset=True, horizon=horizon, philox_seed=seed,
                   env_id_offset=12345, **materialise(copy.deepcopy(cfg)))
    env.set_jit(jit)
    first = ora.reset()  # mirrors the reset at the end of the env constructor
    cur = (np.stack([env._cur.cpu().numpy(), env._cur_irr.cpu().numpy()], -1)
           if irr else env._cur.cpu().numpy())
    assert np.array_equal(cur, first)
    r = np.random.default_rng(seed)
    for part, given in ((T, False), (7, True), (1, True)):
        actions = None
        if given:
            actions = r.integers(0, ora.A, size=(part, N))
            if irr:
                actions = np.stack([actions, r.integers(0, ora.A1, size=(part, N))], -1)
        want = ora.rollout(part, actions=actions)
        got = env.rollout(part, actions=actions)
        for k in ("obs", "final_obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, cfg)
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12, err_msg=str(cfg))
    # the standard signature of the bench (actions and out= given, no final_obs):
    # the FAST kernel with its action-prefetch loop
    part = 21
    actions = r.integers(0, ora.A, size=(part, N))
    if irr:
        actions = np.stack([actions, r.integers(0, ora.A1, size=(part, N))], -1)
    want = ora.rollout(part, actions=actions)
    acts_dev = torch.as_tensor(actions, dtype=torch.int32, device="cuda")
    out = env.rollout(part, actions=acts_dev, want_final_obs=False)
    for k in ("obs", "terminated", "truncated"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), (k, cfg)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-12, atol=1e-12, err_msg=str(cfg))
    st = env.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert st[k][0] == ora.stats[k], (k, cfg)


@pytest.mark.parametrize("seed", CONTINUOUS_SEEDS)
def test_continuous_fuzz_cuda_vs_oracle(seed):
    cfg = continuous_fuzz_config(seed)
    N, T, horizon, jit = _shape_of(seed)
    ora = VectorContinuousOracle(scalar_oracle(materialise(copy.deepcopy(cfg))), N,
                                 autoreset=True, horizon=horizon, seed=seed,
                                 env_id_offset=777)
    env = make_env(N, autoreset=True, horizon=horizon, philox_seed=seed,
                   env_id_offset=777, **materialise(copy.deepcopy(cfg)))
    env.set_jit(jit)
    ora.reset()
    np.testing.assert_allclose(env.curr_obs.cpu().numpy(), ora.em, rtol=1e-6)
    D, amax = cfg["state_space_dim"], cfg["action_space_max"]
    # a few percent of the actions leave the action space (the reference
    # freezes the state for those, rl_toy_env.py:1640-1679)
    acts = np.random.default_rng(seed).uniform(
        -1.03 * amax, 1.03 * amax, size=(T, N, D)).astype(np.float32)
    want = ora.rollout(T, acts)
    got = env.rollout(T, torch.as_tensor(acts))
    for k in ("terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, cfg)
    for k in ("obs", "final_obs"):
        np.testing.assert_allclose(got[k].cpu().numpy(), want[k], rtol=1e-5,
                                   atol=1e-6, err_msg=str(cfg))
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-4, atol=1e-5, err_msg=str(cfg))


@pytest.mark.parametrize("seed", GRID_SEEDS)
def test_grid_fuzz_cuda_vs_oracle(seed):
    cfg = grid_fuzz_config(seed)
    N, T, horizon, _ = _shape_of(seed)
    ora = VectorGridOracle(scalar_oracle(materialise(copy.deepcopy(cfg))), N,
                           autoreset=True, horizon=horizon, seed=seed,
                           env_id_offset=99)
    env = make_env(N, autoreset=True, horizon=horizon, philox_seed=seed,
                   env_id_offset=99, **materialise(copy.deepcopy(cfg)))
    first = ora.reset()
    assert np.array_equal(env.get_augmented_state()["curr_state"].cpu().numpy(), first)
    rng = np.random.default_rng(seed)
    for part in (T, 1):
        acts = np.zeros((part, N, ora.nd), dtype=np.int64)
        d = rng.integers(ora.nd, size=(part, N))
        np.put_along_axis(acts, d[..., None], rng.integers(-1, 2, size=(part, N, 1)), -1)
        bad = rng.random((part, N)) < 0.05  # not unit moves: no-ops in the reference
        acts[bad] = rng.integers(-2, 3, size=(int(bad.sum()), ora.nd))
        want = ora.rollout(part, acts)
        got = env.rollout(part, actions=acts)
        for k in ("obs", "final_obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, cfg)
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12, err_msg=str(cfg))
    st = env.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert st[k][0] == ora.stats[k], (k, cfg)


@pytest.mark.parametrize("seed", IMAGE_SEEDS)
def test_image_fuzz_same_seed_drop_in(seed):
    """noise='numpy': one env with the configuration and seed of the (oracle of
    the) reference consumes the reference's own numpy streams => the same
    trajectory and the same pixels, constructor included."""
    cfg = image_fuzz_config(seed)
    cont = cfg["state_space_type"] == "continuous"
    ref = scalar_oracle(materialise(copy.deepcopy(cfg)))
    env = make_env(1, noise="numpy", **materialise(copy.deepcopy(cfg)))
    assert np.array_equal(env.curr_obs[0].cpu().numpy(), ref.curr_obs)
    rng = np.random.default_rng(seed)
    irr = (not cont) and bool(cfg.get("irrelevant_features"))
    for t in range(30):
        if cont:
            a = rng.uniform(-1, 1, size=cfg["state_space_dim"]).astype(np.float32)
            o2, r2, d2, _, _ = env.step(torch.as_tensor(a.copy())[None])
        elif irr:
            a = [int(rng.integers(n)) for n in cfg["action_space_size"]]
            o2, r2, d2, _, _ = env.step([a])
        else:
            a = int(rng.integers(cfg["action_space_size"]))
            o2, r2, d2, _, _ = env.step([a])
        o1, r1, d1, _, _ = ref.step(a)
        assert np.array_equal(o2[0].cpu().numpy(), o1), (t, cfg)
        assert bool(d2[0]) == d1
        if cont:
            np.testing.assert_allclose(float(r2[0]), float(r1), rtol=1e-5, atol=1e-6)
        else:
            assert float(r2[0]) == float(r1)
        if d1 or t % 10 == 9:
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert np.array_equal(o2[0].cpu().numpy(), o1), (t, cfg)


def _synthetic_wrapper_record(spec, K, T, seed):
    """Inputs of the GymEnvWrapper tail in the layout of the wrap_*.npz
    records (what make_wrapper_golden records from the reference), drawn from a
    seeded generator instead: the GPU box has no reference to record from."""
    from oracle.wrapper_tail import ScalarWrapperTail
    r = np.random.default_rng(seed)
    cfg = spec["config"]
    cont = cfg["state_space_type"] == "continuous"
    image = spec["base"] == "image"
    dim = spec.get("dim", 1)
    g = {}
    if cont:
        g["action"] = r.uniform(-1, 1, (K, T, dim)).astype(np.float32)
        g["base_obs"] = r.uniform(-3, 3, (K, T, dim)).astype(np.float32)
    else:
        g["action"] = r.integers(spec["n_actions"], size=(K, T))
        if image:
            s = spec["side"]
            g["base_obs"] = r.integers(0, 256, (K, T, s, s, 3)).astype(np.uint8)
        else:
            g["base_obs"] = r.integers(50, size=(K, T))
    pn = cfg.get("transition_noise")
    g["choice_u"] = r.random((K, T)) if (pn and not cont) else np.full((K, T), np.nan)
    g["obs_noise"] = (r.normal(0, pn, (K, T, dim)) if (pn and cont)
                      else np.full((K, T, dim), np.nan))
    g["reward_noise"] = (r.normal(0, cfg["reward_noise"], (K, T))
                         if "reward_noise" in cfg else np.full((K, T), np.nan))
    if image:
        lo, hi = ScalarWrapperTail(n_actions=spec["n_actions"], **cfg).shift_draw_bounds(
            spec["side"])
        g["shift"] = r.integers(lo, hi, size=(K, T, 2))
    else:
        g["shift"] = np.zeros((K, T, 2), dtype=np.int64)
    g["base_reward"] = np.round(r.normal(size=(K, T)), 3)
    g["base_done"] = r.random((K, T)) < 0.08
    return g


@pytest.mark.parametrize("seed", __import__("tests.fuzz_configs", fromlist=["x"]).WRAPPER_SEEDS)
def test_wrapper_tail_fuzz_cuda_vs_oracle(seed):
    """VectorGymEnvTail (replayed draws) against the oracle's tail on the
    configurations tests/test_fuzz_reference.py pins the oracle on against the
    reference wrapper; terminal steps included (the reference raises there,
    gym_env_wrapper.py:414, so they are CUDA-vs-oracle only)."""
    from tests.fuzz_configs import wrapper_fuzz_spec
    from tests.test_wrapper_tail import CudaLanes, OracleLanes
    spec = wrapper_fuzz_spec(seed)
    K, T = 37, 40
    g = _synthetic_wrapper_record(spec, K, T, seed)
    ora, dev = OracleLanes(K, spec), CudaLanes(K, spec)
    for t in range(T):
        a1 = ora.actions(g["action"][:, t], g["choice_u"][:, t])
        a2 = dev.actions(g["action"][:, t], g["choice_u"][:, t])
        assert np.array_equal(np.asarray(a1), np.asarray(a2)), (t, spec)
        args = (g["base_obs"][:, t], g["base_reward"][:, t], g["base_done"][:, t],
                g["reward_noise"][:, t], g["obs_noise"][:, t], g["shift"][:, t])
        o1, r1 = ora.post(*args)
        o2, r2 = dev.post(*args)
        assert np.array_equal(np.asarray(r1), np.asarray(r2)), (t, spec)
        assert np.array_equal(np.asarray(o1), np.asarray(o2)), (t, spec)
