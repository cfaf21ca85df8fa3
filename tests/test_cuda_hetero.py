"""GPU tests of heterogeneous config groups (BASELINE config #5 shape) and of
the shard-aware global env ids: a multi-group launch equals the per-group
single-config envs, and a job split over 2 shards equals the unsplit job."""
import warnings

import numpy as np
import pytest
import torch

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def make_env(*a, **k):
    from mdp_playground_b200 import VectorRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorRLToyEnv(*a, **k)


NAMES = ["c1_seq1", "c2_every1", "big50", "custom_8x5", "diam3_seq4",
         "seq2_denser", "notmax_diam2"]
SIZES = [300, 500, 200, 100, 257, 64, 131]


def groups():
    return [gu.case_config(n) for n in NAMES]


def test_multi_group_launch_equals_single_group_envs():
    N, T = sum(SIZES), 40
    het = make_env(N, autoreset=True, horizon=9, philox_seed=5,
                   config_groups=groups(), group_sizes=SIZES)
    out = het.rollout(T, want_final_obs=True)
    begin = 0
    for cfg, n, sl in zip(groups(), SIZES, het.group_slices):
        one = make_env(n, autoreset=True, horizon=9, philox_seed=5,
                       env_id_offset=begin, **cfg)
        ref = one.rollout(T, want_final_obs=True)
        for k in ref:
            assert torch.equal(out[k][:, sl], ref[k]), (k, begin)
        begin += n
    st = het.episode_stats()
    assert st["transitions"].tolist() == [n * T for n in SIZES]
    assert st["episodes"].sum() == int((out["terminated"] | out["truncated"]).sum())


def test_multi_group_jit_equals_aot():
    """Partially specialised NVRTC kernel (scalars shared by all groups are
    literals) vs the ahead-of-time kernel: bit-identical."""
    N, T = sum(SIZES), 48
    outs = []
    for jit in (True, False):
        env = make_env(N, autoreset=True, horizon=11, philox_seed=8,
                       config_groups=groups(), group_sizes=SIZES)
        env.set_jit(jit)
        acts = torch.randint(0, 5, (T, N), dtype=torch.int32, device="cuda",
                             generator=torch.Generator("cuda").manual_seed(2))
        outs.append(env.rollout(T, actions=acts, want_final_obs=False))
        assert env.jit_last_used == jit, env.jit_log
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_two_shards_equal_the_unsplit_job():
    T = 25
    half = [s // 2 for s in SIZES]
    full = [2 * h for h in half]
    whole = make_env(sum(full), autoreset=True, horizon=7, philox_seed=3,
                     config_groups=groups(), group_sizes=full)
    r0 = make_env(sum(half), autoreset=True, horizon=7, philox_seed=3,
                  config_groups=groups(), group_sizes=half, shard=(0, 2))
    r1 = make_env(sum(half), autoreset=True, horizon=7, philox_seed=3,
                  config_groups=groups(), group_sizes=half, shard=(1, 2))
    ow, o0, o1 = whole.rollout(T), r0.rollout(T), r1.rollout(T)
    for k in ow:
        for sw, s0 in zip(whole.group_slices, r0.group_slices):
            want = ow[k][:, sw]
            got = torch.cat([o0[k][:, s0], o1[k][:, s0]], dim=1)
            assert torch.equal(want, got), k
    s = whole.episode_stats()
    a, b = r0.episode_stats(), r1.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert np.array_equal(s[k], a[k] + b[k]), k
    np.testing.assert_allclose(s["reward"], a["reward"] + b["reward"], rtol=1e-12)


def test_sweep_grid_smoke_1000_groups():
    """The C5 grid: delay x seq_len x P-noise x R-noise x make_denser."""
    base = gu.case_config("c1_seq1")
    cfgs = []
    for d in (0, 1, 2, 4, 8):
        for L in (1, 2, 3, 4):
            for pn in (0, 0.01, 0.02, 0.1, 0.25):
                for rn in (0, 1, 5, 10, 25):
                    for md in (False, True):
                        cfgs.append(dict(base, delay=d, sequence_length=L,
                                         transition_noise=pn, reward_noise=rn,
                                         make_denser=md,
                                         reward_every_n_steps=True))
    assert len(cfgs) == 1000
    N = 64 * 1000
    env = make_env(N, autoreset=True, horizon=100, config_groups=cfgs)
    out = env.rollout(50, want_final_obs=False)
    st = env.episode_stats()
    assert np.all(st["transitions"] == 64 * 50)
    # noise-free groups draw nothing, noisy ones do
    pn = np.array([c["transition_noise"] for c in cfgs])
    assert np.all(st["noisy_transitions"][pn == 0] == 0)
    assert np.all(st["noisy_transitions"][pn == 0.25] > 0)
    rn = np.array([c["reward_noise"] for c in cfgs])
    assert np.all(st["abs_reward_noise"][rn == 0] == 0)
    ratio = st["abs_reward_noise"][rn == 25].mean() / st["abs_reward_noise"][rn == 5].mean()
    assert abs(ratio - 5) < 0.3
    assert int(out["obs"].max()) < 8


def _mixed_groups():
    """40 cells of the BASELINE config #5 grid (delay x sequence_length x noise
    x make_denser; experiments/dqn_seq_del.py:7-20, dqn_p_r_noises.py:7-20)."""
    base = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
                state_space_size=8, action_space_size=8, reward_density=0.25,
                terminal_state_density=0.25, reward_every_n_steps=True)
    rng = np.random.default_rng(40)
    cells = [(d, L, pn, rn, md) for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4)
             for pn in (0, 0.01, 0.1, 0.25) for rn in (0, 1, 5, 25)
             for md in (False, True)]
    pick = rng.choice(len(cells), size=40, replace=False)
    return [dict(base, delay=cells[i][0], sequence_length=cells[i][1],
                 transition_noise=cells[i][2], reward_noise=cells[i][3],
                 make_denser=cells[i][4]) for i in pick]


@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("normal", ["fp64", "fast"])
def test_mixed_40_group_launch_matches_grouped_oracle(jit, normal):
    """ONE heterogeneous launch against the CPU oracle directly (not against
    other CUDA envs): 40 mixed groups of 37 envs (partially filled CTAs),
    given actions, auto-reset + horizon; T = 75 covers two staged ziggurat
    windows + a chunk drawn directly.  States / flags bit-exact, rewards 1e-12
    (1e-5 x sigma for the fp32 SFU normals)."""
    from oracle.scalar_env import ScalarRLToyEnv
    from oracle.vector_oracle import VectorGroupedOracle
    cfgs = _mixed_groups()
    sizes = [37] * len(cfgs)
    N, T = sum(sizes), 75
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalars = [ScalarRLToyEnv(**dict(c)) for c in cfgs]
    ora = VectorGroupedOracle(scalars, sizes, autoreset=True, horizon=23, seed=3,
                              env_id_offset=1000, fast_normal=normal == "fast")
    ora.reset()
    env = make_env(N, autoreset=True, horizon=23, philox_seed=3, env_id_offset=1000,
                   config_groups=[dict(c) for c in cfgs], group_sizes=sizes,
                   normal_precision=normal)
    env.set_jit(jit)
    acts = np.random.default_rng(1).integers(0, 8, size=(T, N))
    got = env.rollout(T, actions=torch.as_tensor(acts, dtype=torch.int32, device="cuda"))
    assert env.jit_last_used == jit, env.jit_log
    want = ora.rollout(T, actions=acts)
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    tol = 1e-5 * 25 if normal == "fast" else 1e-12
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=0 if normal == "fast" else 1e-12, atol=tol)
    st = env.episode_stats()
    for g, s in enumerate(ora.stats):
        for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
            assert st[k][g] == s[k], (g, k)


def test_mixed_40_group_long_launch_standard_signature():
    """The same 40 groups through the standard signature (no final_obs: the
    FAST multi-group kernel, chunks of 4, 16-step ziggurat windows) with T = 150,
    i.e. nine windows and a partial one, against the grouped oracle and against
    the ahead-of-time kernel."""
    from oracle.scalar_env import ScalarRLToyEnv
    from oracle.vector_oracle import VectorGroupedOracle
    cfgs = _mixed_groups()
    sizes = [37] * len(cfgs)
    N, T = sum(sizes), 150
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalars = [ScalarRLToyEnv(**dict(c)) for c in cfgs]
    ora = VectorGroupedOracle(scalars, sizes, autoreset=True, horizon=23, seed=3,
                              env_id_offset=1000)
    ora.reset()
    acts = np.random.default_rng(1).integers(0, 8, size=(T, N))
    want = ora.rollout(T, actions=acts)
    rewards = []
    for jit in (True, False):
        env = make_env(N, autoreset=True, horizon=23, philox_seed=3, env_id_offset=1000,
                       config_groups=[dict(c) for c in cfgs], group_sizes=sizes)
        env.set_jit(jit)
        got = env.rollout(T, actions=torch.as_tensor(acts, dtype=torch.int32, device="cuda"),
                          want_final_obs=False)
        assert env.jit_last_used == jit, env.jit_log
        for k in ("obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (jit, k)
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12)
        rewards.append(got["reward"])
    assert torch.equal(rewards[0], rewards[1])


def _continuous_cells():
    """A sac_move_to_a_point_* style grid: time_unit x action_space_max x noise
    x delay x target_radius cells of one continuous env (same dim / order)."""
    base = dict(seed=0, state_space_type="continuous", action_space_type="continuous",
                state_space_dim=4, action_space_dim=4, relevant_indices=[0, 1],
                irrelevant_features=True, transition_dynamics_order=2, inertia=1.0,
                target_point=[0.0, 0.0], state_space_max=4.0, make_denser=True)
    cells = []
    for tu, amax, pn, rn, d, rad, alw, order in [
            (0.1, 1.0, 0, 0, 0, 0.5, 0.0, 2), (0.5, 0.25, 0.05, 0, 2, 0.5, 0.0, 1),
            (1.0, 0.5, 0, 0.2, 1, 1.0, 0.0, 2), (0.25, 1.0, 0.1, 0.1, 4, 0.25, 0.5, 3),
            (0.05, 0.1, 0, 0, 0, 0.5, 0.0, 2), (0.5, 1.0, 0.02, 0.3, 3, 2.0, 0.0, 1)]:
        c = dict(base, time_unit=tu, action_space_max=amax, delay=d,
                 target_radius=rad, action_loss_weight=alw, reward_scale=1.0 + d,
                 transition_dynamics_order=order)
        if pn:
            c["transition_noise"] = pn
        if rn:
            c["reward_noise"] = rn
        cells.append(c)
    return cells


def test_continuous_multi_group_launch_equals_single_group_envs_and_oracle():
    """config_groups for continuous envs (one launch per sweep): the
    heterogeneous launch == one env per cell (CUDA, bit-exact) == the grouped
    CPU oracle (1e-5, the continuous contract); 2 shards == unsplit."""
    from oracle.scalar_env import ScalarRLToyEnv
    from oracle.vector_continuous_oracle import VectorGroupedContinuousOracle
    cells = _continuous_cells()
    sizes = [150, 130, 128, 64, 200, 97]
    N, T = sum(sizes), 40
    acts = (torch.rand((T, N, 4), device="cuda",
                       generator=torch.Generator("cuda").manual_seed(6)) * 2.2 - 1.1)
    het = make_env(N, autoreset=True, horizon=17, philox_seed=9,
                   config_groups=[dict(c) for c in cells], group_sizes=sizes)
    obs0 = het.curr_obs.clone()
    out = het.rollout(T, actions=acts)
    begin = 0
    for c, n, sl in zip(cells, sizes, het.group_slices):
        one = make_env(n, autoreset=True, horizon=17, philox_seed=9,
                       env_id_offset=begin, **dict(c))
        one.set_jit(False)   # (the group launch runs the ahead-of-time kernel)
        assert torch.equal(one.curr_obs, obs0[sl])
        ref = one.rollout(T, actions=acts[:, sl].contiguous())
        for k in ref:
            assert torch.equal(out[k][:, sl], ref[k]), (k, begin)
        begin += n
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalars = [ScalarRLToyEnv(**dict(c)) for c in cells]
    ora = VectorGroupedContinuousOracle(scalars, sizes, autoreset=True, horizon=17, seed=9)
    ora.reset()
    want = ora.rollout(T, acts.cpu().numpy())
    for k in ("terminated", "truncated"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(out["obs"].cpu().numpy(), want["obs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-4, atol=1e-5)
    st = het.episode_stats()
    assert st["transitions"].tolist() == [n * T for n in sizes]
    # two shards of the same job
    half = [n // 2 for n in [128, 128, 256, 64, 192, 128]]
    whole = make_env(2 * sum(half), autoreset=True, horizon=17, philox_seed=9,
                     config_groups=[dict(c) for c in cells],
                     group_sizes=[2 * n for n in half])
    a2 = (torch.rand((T, 2 * sum(half), 4), device="cuda") * 2 - 1)
    ref = whole.rollout(T, actions=a2)
    for r in range(2):
        part = make_env(sum(half), autoreset=True, horizon=17, philox_seed=9,
                        config_groups=[dict(c) for c in cells], group_sizes=half,
                        shard=(r, 2))
        cols = torch.cat([torch.arange(sl.start + r * n, sl.start + (r + 1) * n)
                          for sl, n in zip(whole.group_slices, half)]).cuda()
        got = part.rollout(T, actions=a2[:, cols].contiguous())
        for k in ref:
            assert torch.equal(got[k], ref[k][:, cols]), (k, r)


def _grid_cells():
    base = dict(seed=0, state_space_type="grid", delay=0, sequence_length=1,
                reward_function="move_to_a_point")
    return [
        dict(base, grid_shape=(8, 8), target_point=[5, 5], make_denser=True,
             reward_scale=3.0, term_state_reward=-0.25, terminal_states=[[5, 5]]),
        dict(base, grid_shape=(8, 8), target_point=[5, 5], make_denser=False,
             transition_noise=0.3, reward_noise=1.0, reward_scale=2.0,
             reward_shift=0.5, term_state_reward=2.0),
        dict(base, grid_shape=(5, 9), target_point=[1, 7], make_denser=True,
             transition_noise=0.2, reward_every_n_steps=2),
        dict(base, grid_shape=(3, 4), target_point=[0, 0], make_denser=False,
             reward_noise=0.5),
        dict(base, grid_shape=(12, 6), target_point=[11, 2], make_denser=True,
             transition_noise=1.0, reward_shift=-1.0),
    ]


def test_grid_multi_group_launch_equals_single_group_envs_and_oracle():
    """config_groups for grid envs (one launch per sweep): the heterogeneous
    launch == one env per cell (CUDA, bit-exact) == the grouped CPU oracle
    (cells exact, rewards 1e-12); per-group statistics; 2 shards == unsplit."""
    from oracle.scalar_env import ScalarRLToyEnv
    from oracle.vector_grid_oracle import VectorGroupedGridOracle
    cells = _grid_cells()
    sizes = [150, 130, 128, 64, 200]
    N, T = sum(sizes), 40
    gen = torch.Generator("cuda").manual_seed(6)
    acts = torch.zeros((T, N, 2), dtype=torch.int64, device="cuda")
    dim = torch.randint(0, 2, (T, N, 1), device="cuda", generator=gen)
    acts.scatter_(2, dim, torch.randint(-1, 2, (T, N, 1), device="cuda", generator=gen))
    bad = torch.rand((T, N), device="cuda", generator=gen) < 0.05   # invalid: no-ops
    acts[bad] = torch.randint(-2, 3, (int(bad.sum()), 2), device="cuda", generator=gen)
    het = make_env(N, autoreset=True, horizon=17, philox_seed=9,
                   config_groups=[dict(c) for c in cells], group_sizes=sizes)
    assert het.n_groups == 5
    obs0 = het.curr_obs.clone()
    out = het.rollout(T, actions=acts, want_final_obs=True)
    begin = 0
    for c, n, sl in zip(cells, sizes, het.group_slices):
        one = make_env(n, autoreset=True, horizon=17, philox_seed=9,
                       env_id_offset=begin, **dict(c))
        assert torch.equal(one.curr_obs, obs0[sl])
        ref = one.rollout(T, actions=acts[:, sl].contiguous(), want_final_obs=True)
        for k in ref:
            assert torch.equal(out[k][:, sl], ref[k]), (k, begin)
        begin += n
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scalars = [ScalarRLToyEnv(**dict(c)) for c in cells]
    ora = VectorGroupedGridOracle(scalars, sizes, autoreset=True, horizon=17, seed=9)
    assert np.array_equal(ora.reset(), obs0.cpu().numpy())
    want = ora.rollout(T, acts.cpu().numpy())
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-12, atol=1e-12)
    st = het.episode_stats()
    assert st["transitions"].tolist() == [n * T for n in sizes]
    for k in ("episodes", "noisy_transitions", "terminated"):
        assert st[k].tolist() == [p.stats[k] for p in ora.parts], k
    # the fast path (no final_obs) and single steps run the same kernels
    het2 = make_env(N, autoreset=True, horizon=17, philox_seed=9,
                    config_groups=[dict(c) for c in cells], group_sizes=sizes)
    for t in range(5):
        o, r, te, tr, _ = het2.step(acts[t])
        assert torch.equal(o, out["obs"][t]) and torch.equal(r, out["reward"][t])
        assert torch.equal(te, out["terminated"][t]) and torch.equal(tr, out["truncated"][t])
    # two shards of the same job
    half = [64, 64, 128, 32, 96]
    whole = make_env(2 * sum(half), autoreset=True, horizon=17, philox_seed=9,
                     config_groups=[dict(c) for c in cells],
                     group_sizes=[2 * n for n in half])
    a2 = torch.zeros((T, 2 * sum(half), 2), dtype=torch.int64, device="cuda")
    a2[..., 1] = torch.randint(-1, 2, (T, 2 * sum(half)), device="cuda", generator=gen)
    ref = whole.rollout(T, actions=a2)
    for r in range(2):
        part = make_env(sum(half), autoreset=True, horizon=17, philox_seed=9,
                        config_groups=[dict(c) for c in cells], group_sizes=half,
                        shard=(r, 2))
        cols = torch.cat([torch.arange(sl.start + r * n, sl.start + (r + 1) * n)
                          for sl, n in zip(whole.group_slices, half)]).cuda()
        got = part.rollout(T, actions=a2[:, cols].contiguous())
        for k in ref:
            assert torch.equal(got[k], ref[k][:, cols]), (k, r)


def test_grid_groups_must_agree_on_the_number_of_coordinates():
    cells = _grid_cells()[:2]
    cells[1] = dict(cells[1], irrelevant_features=True)
    with pytest.raises(ValueError, match="number of grid dimensions"):
        make_env(64, config_groups=cells)
