"""Discrete `irrelevant_features` (rl_toy_env.py:1154-1230 tables, :2029-2035
and :2062-2088 step, :2255-2264 reset): a second,
reward-free sub-MDP; states, actions and observations are rows (relevant,
irrelevant).  CPU tests pin the host tables and the batched oracle to goldens
recorded from the unmodified reference; GPU tests run the CUDA path against
those goldens, the scalar oracle (same-seed drop-in) and the batched oracle
(native Philox noise).  All comparisons are bit-exact except Philox-mode
rewards (1e-12: device log / sincospi vs numpy)."""
import json
import warnings

import numpy as np
import pytest

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_oracle import VectorDiscreteOracle
from tests import golden_util as gu
from tests.golden.cases import CASES

IRR_PLAIN = [n for n in gu.IRR_CASES
             if not CASES[n]["config"].get("image_representations")]


def scalar_oracle(cfg):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ScalarRLToyEnv(**cfg)


def make_env(*a, **k):
    from mdp_playground_b200 import VectorRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorRLToyEnv(*a, **k)


def replay_irr_golden(vec_reset, vec_step, g):
    K, T = g["done"].shape
    cur = vec_reset(None, np.stack([g["init_reset_u"], g["init_irr_reset_u"]], -1))
    assert np.array_equal(cur, g["init_state"])
    for t in range(T):
        obs, r, done = vec_step(
            g["actions"][:, t], np.nan_to_num(g["transition_u"][:, t]),
            np.nan_to_num(g["irr_transition_u"][:, t]),
            np.nan_to_num(g["reward_noise"][:, t]), t)
        assert np.array_equal(obs, g["state"][:, t]), t
        assert np.array_equal(r, g["reward"][:, t]), t
        assert np.array_equal(done, g["done"][:, t]), t
        m = g["reset_after"][:, t]
        if m.any():
            ru = np.nan_to_num(np.stack([g["reset_u"][:, t], g["irr_reset_u"][:, t]], -1))
            cur = vec_reset(m, ru, t)
            assert np.array_equal(cur[m], g["reset_state"][m, t]), t


@pytest.mark.parametrize("name", gu.IRR_CASES)
def test_host_tables_equal_reference_golden(name):
    from mdp_playground_b200.config import parse_config
    from mdp_playground_b200.tables import build_discrete_tables
    g = gu.load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = parse_config(gu.case_config(name))
        tb = build_discrete_tables(sp)
    assert sp.seed_dict == json.loads(str(g["seed_dict"]))
    assert np.array_equal(tb.transition, g["P"])
    assert np.array_equal(tb.transition_irr, g["P_irr"])
    assert list(tb.rewardable_sequences.items()) == list(gu.golden_sequences(g).items())
    S1 = tb.n_states_irr
    assert np.allclose(tb.init_cdf_irr, np.arange(1, S1 + 1) / S1)


@pytest.mark.parametrize("name", IRR_PLAIN)
def test_vector_oracle_replays_reference_golden(name):
    g = gu.load(name)
    vec = VectorDiscreteOracle(scalar_oracle(gu.case_config(name)), g["done"].shape[0])

    def vec_reset(mask, reset_u, t=None):
        return vec.reset(mask=mask, reset_u=reset_u)

    def vec_step(a, u, u1, n, t):
        out = vec.rollout(1, actions=a[None], replay=dict(
            transition_u=u[None], irr_transition_u=u1[None], reward_noise=n[None]))
        return out["obs"][0], out["reward"][0], out["terminated"][0]

    replay_irr_golden(vec_reset, vec_step, g)


# --------------------------------------------------------------------------
# GPU
# --------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.IRR_CASES)
def test_cuda_replays_reference_golden(name):
    """Recorded reference draws (both sub-spaces, and the transform draws of
    both sub-images) replayed through the kernels: bit-exact."""
    g = gu.load(name)
    K = g["done"].shape[0]
    image = "obs_image" in g
    env = make_env(K, noise="replay", **gu.case_config(name))
    assert np.array_equal(env.transition_matrix, g["P"])
    assert np.array_equal(env.transition_matrix_irrelevant, g["P_irr"])

    def vec_reset(mask, reset_u, t=None):
        opt = {"mask": mask, "reset_u": reset_u}
        if image:
            opt["image_params"] = g["init_image_params"] if t is None else np.where(
                mask[:, None, None], g["reset_image_params"][:, t],
                g["image_params"][:, t])
        obs, _ = env.reset(options=opt)
        if image:
            want = g["init_image"] if t is None else g["reset_image"][:, t]
            sel = slice(None) if mask is None else mask
            assert np.array_equal(obs.cpu().numpy()[sel], want[sel])
        return env.get_augmented_state()["curr_state"].cpu().numpy()

    def vec_step(a, u, u1, n, t):
        rep = dict(transition_u=u, irr_transition_u=u1, reward_noise=n)
        if image:
            rep["image_params"] = g["image_params"][:, t]
        obs, r, term, trunc, info = env.step(a, replay=rep)
        assert not trunc.any()
        if image:
            assert obs.shape[1:] == (200, 100, 1)
            assert np.array_equal(obs.cpu().numpy(), g["obs_image"][:, t]), t
        return info["state"].cpu().numpy(), r.cpu().numpy(), term.cpu().numpy()

    replay_irr_golden(vec_reset, vec_step, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.IRR_CASES)
def test_same_seed_drop_in_numpy_streams(name):
    """noise='numpy': one env, same config and seed as the (oracle of the)
    reference => the same trajectory, constructor included."""
    cfg = gu.case_config(name)
    ref = scalar_oracle(gu.case_config(name))
    env = make_env(1, noise="numpy", **cfg)
    image = bool(cfg.get("image_representations"))

    def same(o_env, o_ref):
        return np.array_equal(o_env[0].cpu().numpy(), np.asarray(o_ref))
    assert same(env.curr_obs, ref.curr_obs)
    rng = np.random.default_rng(9)
    A = cfg["action_space_size"]
    for t in range(40 if image else 150):
        a = np.array([rng.integers(A[0]), rng.integers(A[1])])
        o1, r1, d1, _, _ = ref.step(a)
        o2, r2, d2, tr2, info = env.step(a[None])
        assert same(o2, o1) and float(r2[0]) == float(r1), t
        assert np.array_equal(info["state"][0].cpu().numpy(), ref.curr_state)
        assert bool(d2[0]) == d1
        if d1 or t % 20 == 19:
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert same(o2, o1)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,T,autoreset,horizon", [
    ("irr_8x8_noise", 1500, 40, True, 9),
    ("irr_6x10_diam2", 700, 33, True, 0),
    ("irr_8x5_det", 300, 30, False, 0),
])
@pytest.mark.parametrize("jit", [True, False])
def test_philox_rollout_matches_oracle(name, N, T, autoreset, horizon, jit):
    """Native Philox noise, device-drawn and given actions, the specialised and
    the ahead-of-time kernel: states of both chains bit-exact."""
    cfg = gu.case_config(name)
    ora = VectorDiscreteOracle(scalar_oracle(gu.case_config(name)), N,
                               autoreset=autoreset, horizon=horizon, seed=77,
                               env_id_offset=1000)
    env = make_env(N, autoreset=autoreset, horizon=horizon, philox_seed=77,
                   env_id_offset=1000, **cfg)
    env.set_jit(jit)
    first = ora.reset()  # mirrors the reset at the end of the env constructor
    assert np.array_equal(env.get_augmented_state()["curr_state"].cpu().numpy()
                          if env.track_history else np.stack(
                              [env._cur.cpu().numpy(), env._cur_irr.cpu().numpy()], -1),
                          first)
    for part, acts in ((T, None), (7, "given"), (1, "given")):
        actions = None
        if acts:
            r = np.random.default_rng(3)
            actions = np.stack([r.integers(0, ora.A, size=(part, N)),
                                r.integers(0, ora.A1, size=(part, N))], -1)
        want = ora.rollout(part, actions=actions)
        got = env.rollout(part, actions=actions)
        assert env.jit_last_used == jit, env.jit_log
        assert got["obs"].shape == (part, N, 2)
        for k in ("obs", "final_obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12)
    assert (want["obs"][..., 1] != want["obs"][..., 0]).any()


@pytest.mark.gpu
def test_graphed_step_and_fast_signature_with_irrelevant_features():
    """CUDA-graph step == eager step; the standard-signature (no final_obs)
    specialised kernel == the generic one."""
    import torch
    cfg = gu.case_config("irr_8x8_noise")
    a = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    b = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    step = b.make_graphed_step()
    gen = torch.Generator("cuda").manual_seed(0)
    for t in range(12):
        acts = torch.randint(0, 8, (512, 2), dtype=torch.int32, device="cuda",
                             generator=gen)
        o1, r1, d1, t1, _ = a.step(acts)
        o2, r2, d2, t2, _ = step(acts)
        assert torch.equal(o1, o2) and torch.equal(r1, r2)
        assert torch.equal(d1, d2) and torch.equal(t1, t2)
    c = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    d = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    acts = torch.randint(0, 8, (64, 512, 2), dtype=torch.int32, device="cuda",
                         generator=gen)
    x = c.rollout(64, actions=acts, want_final_obs=False)
    y = d.rollout(64, actions=acts, want_final_obs=True)
    for k in x:
        assert torch.equal(x[k], y[k]), k
    assert torch.equal(y["final_obs"][~(y["terminated"] | y["truncated"])],
                       y["obs"][~(y["terminated"] | y["truncated"])])


@pytest.mark.gpu
@pytest.mark.parametrize("jit", [True, False])
def test_multi_group_launch_with_irrelevant_features(jit):
    """Heterogeneous groups that all carry an irrelevant sub-MDP (different
    sizes / noise / delay per group): one launch == the per-group envs; mixing
    groups with and without one is rejected."""
    import torch
    names = ["irr_8x8_noise", "irr_6x10_diam2", "irr_8x5_det"]
    sizes = [300, 129, 64]
    cfgs = [gu.case_config(n) for n in names]
    het = make_env(sum(sizes), autoreset=True, horizon=9, philox_seed=5,
                   config_groups=cfgs, group_sizes=sizes)
    het.set_jit(jit)
    out = het.rollout(40, want_final_obs=True)
    assert het.jit_last_used == jit, het.jit_log
    begin = 0
    for name, n, sl in zip(names, sizes, het.group_slices):
        one = make_env(n, autoreset=True, horizon=9, philox_seed=5,
                       env_id_offset=begin, **gu.case_config(name))
        ref = one.rollout(40, want_final_obs=True)
        for k in ref:
            assert torch.equal(out[k][:, sl], ref[k]), (k, name)
        begin += n
    with pytest.raises(AssertionError):
        make_env(128, config_groups=[gu.case_config("irr_8x8_noise"),
                                     gu.case_config("c2_every1")])


@pytest.mark.gpu
def test_full_size_properties_65536_envs():
    """BASELINE-sized batch (65 536 envs x 256 steps, no transition noise, no
    auto-reset): both chains follow their own table -- obs[t] == P[obs[t-1],
    a[t]] per sub-MDP -- and the irrelevant chain never influences rewards or
    termination (same relevant trajectory as an env without it, given the
    same tables)."""
    import torch
    cfg = gu.case_config("irr_8x5_det")
    N, T = 65536, 256
    env = make_env(N, philox_seed=2, track_history=False, **cfg)
    P = torch.as_tensor(env.transition_matrix, device="cuda").long()
    P1 = torch.as_tensor(env.transition_matrix_irrelevant, device="cuda").long()
    start = torch.stack([env._cur, env._cur_irr], -1).long()
    gen = torch.Generator("cuda").manual_seed(0)
    acts = torch.stack([torch.randint(0, 8, (T, N), device="cuda", generator=gen),
                        torch.randint(0, 5, (T, N), device="cuda", generator=gen)],
                       -1).to(torch.int32)
    out = env.rollout(T, actions=acts)
    obs = out["obs"]
    prev = torch.cat([start[None], obs[:-1]], dim=0)
    assert torch.equal(obs[..., 0], P[prev[..., 0], acts[..., 0].long()])
    assert torch.equal(obs[..., 1], P1[prev[..., 1], acts[..., 1].long()])
    term = torch.as_tensor(env.tables.terminal_mask, device="cuda").bool()
    assert torch.equal(out["terminated"], term[obs[..., 0]])
    assert len(torch.unique(obs[..., 1])) == 5
