"""GPU tests of the rarely taken code paths and the awkward sizes: hash
lookup of long sequences, delay rings deeper than the shared-memory limit,
tables too large for shared memory (read from L2), state spaces without the
reset guide table, ragged batch sizes, single env, T = 1."""
import warnings

import numpy as np
import pytest
import torch

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_oracle import VectorDiscreteOracle

pytestmark = pytest.mark.gpu

BASE = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
            reward_density=0.25, terminal_state_density=0.25,
            generate_random_mdp=True, reward_every_n_steps=1)


def make_env(*a, **k):
    from mdp_playground_b200 import VectorRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorRLToyEnv(*a, **k)


def oracle_for(cfg, N, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorDiscreteOracle(ScalarRLToyEnv(**dict(cfg)), N, **kw)


def check(cfg, N, T, jit, autoreset=True, horizon=9, actions=True):
    ora = oracle_for(cfg, N, autoreset=autoreset, horizon=horizon, seed=3)
    env = make_env(N, autoreset=autoreset, horizon=horizon, philox_seed=3,
                   **dict(cfg))
    env.set_jit(jit)
    ora.reset()
    assert np.array_equal(env._cur.cpu().numpy(), ora.cur)
    acts = None
    if actions:
        acts = np.random.default_rng(1).integers(0, ora.A, size=(T, N))
    want = ora.rollout(T, actions=acts)
    got = env.rollout(T, actions=acts)
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-12, atol=1e-12)
    return env


@pytest.mark.parametrize("jit", [True, False])
def test_hash_lookup_long_sequences(jit):
    """16 states x L=4: 16 key bits > 12 -> open-addressing hash in smem."""
    cfg = dict(BASE, state_space_size=16, action_space_size=16,
               sequence_length=4, reward_density=0.02, delay=1,
               transition_noise=0.05, terminal_state_density=0.125)
    env = check(cfg, 700, 60, jit, horizon=25)
    assert len(env.rewardable_sequences) > 300
    assert env.episode_stats()["reward"][0] > 0  # some sequences were hit


@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("denser", [False, True])
def test_byte_indexed_lut_for_12_bit_keys(jit, denser):
    """8 states x L=4: 12 key bits -> an fp64 table would be 32 KB per CTA, so
    the key indexes a byte table of the distinct reward values (LOOKUP_LUT8).
    Delay 3 also exercises the register FIFO of the specialised build."""
    cfg = dict(BASE, state_space_size=8, action_space_size=8, sequence_length=4,
               delay=3, make_denser=denser, transition_noise=0.05,
               terminal_state_density=0.0)
    env = check(cfg, 900, 80, jit, horizon=40)
    assert env.episode_stats()["reward"][0] > 0


@pytest.mark.parametrize("jit", [True, False])
def test_delay_deeper_than_the_shared_memory_ring(jit):
    """delay 20 > 16: the FIFO stays in global memory."""
    cfg = dict(BASE, state_space_size=8, action_space_size=8, sequence_length=1,
               delay=20, reward_noise=0.5, terminal_state_density=0.0)
    env = check(cfg, 300, 70, jit, horizon=45)
    assert env.jit_last_used is False  # deep rings take the AOT kernel


def test_tables_too_large_for_shared_memory():
    """300 states with transition noise: the 300 x 512 fp64 cdf (1.2 MB) and
    its thresholds cannot be staged -> tables are read through L2; 300 > 254
    states also means no reset guide table."""
    cfg = dict(BASE, state_space_size=300, action_space_size=300,
               sequence_length=1, delay=2, transition_noise=0.3,
               reward_density=0.1, terminal_state_density=0.1)
    env = check(cfg, 200, 25, True, horizon=10)
    assert env.jit_last_used is False


@pytest.mark.parametrize("N", [1, 31, 33, 65, 1000])
def test_ragged_batch_sizes(N):
    cfg = dict(BASE, state_space_size=8, action_space_size=8, sequence_length=2,
               delay=1, transition_noise=0.1, reward_noise=0.2)
    check(cfg, N, 23, True)
    check(cfg, N, 1, False, actions=False)


def test_no_autoreset_terminal_states_absorb():
    """Without auto-reset (the reference's behaviour) a terminated env keeps
    returning terminated=True from its absorbing state (:1135-1148)."""
    cfg = dict(BASE, state_space_size=8, action_space_size=8, sequence_length=1)
    env = check(cfg, 500, 30, True, autoreset=False, horizon=0)
    out = env.rollout(5)
    assert out["terminated"].float().mean() > 0.9
    assert not out["truncated"].any()


def test_invalid_arguments_fail_loudly():
    from mdp_playground_b200._lib import MdppError
    cfg = dict(BASE, state_space_size=8, action_space_size=8)
    env = make_env(10, noise="replay", transition_noise=0.1, **cfg)
    with pytest.raises(ValueError):
        env.reset()                       # replay mode without draws
    with pytest.raises(MdppError, match="replay"):
        env.rollout(1, actions=torch.zeros((1, 10), dtype=torch.int32))
    with pytest.raises(AssertionError):
        env.rollout(2, actions=torch.zeros((1, 10), dtype=torch.int32))
    with pytest.raises(KeyError):         # like the reference (:656): no default
        make_env(4, state_space_type="grid", grid_shape=(4, 4))
    with pytest.raises(NotImplementedError):  # crashes in the reference (:1950)
        make_env(4, state_space_type="grid", grid_shape=(4, 4), delay=1,
                 reward_function="move_to_a_point", target_point=[1, 1],
                 make_denser=True)
    c = make_env(4, state_space_type="continuous", state_space_dim=2)
    with pytest.raises(TypeError):        # the reference wants dtype_s actions
        c.step(torch.zeros((4, 2), dtype=torch.float64))


def test_reset_with_the_same_seed_reproduces_the_rollout():
    """gym's reset(seed) contract in Philox mode: the streams are
    re-keyed AND rewound, so reset(seed=s) twice gives the same initial states
    and the same noisy trajectory; a graphed step captured before refuses."""
    import torch
    from tests.test_cuda_discrete import _BENCH_CFG, make_env
    env = make_env(777, autoreset=True, horizon=9, **dict(_BENCH_CFG))
    a = torch.randint(0, 8, (40, 777), dtype=torch.int32, device="cuda")
    gstep = env.make_graphed_step()
    runs = []
    for _ in range(2):
        obs0, _ = env.reset(seed=5)
        out = env.rollout(40, actions=a, want_final_obs=False)
        runs.append((obs0.clone(), {k: v.clone() for k, v in out.items()}))
        env.rollout(3, actions=a[:3], want_final_obs=False)   # (drift in between)
    assert torch.equal(runs[0][0], runs[1][0])
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
    other, _ = env.reset(seed=6)
    assert not torch.equal(other, runs[0][0])
    with pytest.raises(RuntimeError, match="re-seeded"):
        gstep(a[0])


def test_out_of_range_states_are_rejected_host_side():
    import torch
    from tests.test_cuda_discrete import _BENCH_CFG, make_env
    env = make_env(16, track_history=True, **dict(_BENCH_CFG))
    bad = torch.zeros(16, dtype=torch.int64)
    bad[3] = 8
    with pytest.raises(ValueError, match="out of range"):
        env.reset(options={"init_state": bad})
    with pytest.raises(ValueError, match="out of range"):
        env.set_augmented_state(bad.cuda())
    env.reset(options={"init_state": torch.full((16,), 7)})


@pytest.mark.parametrize("kind", ["continuous", "grid"])
def test_two_shards_equal_the_unsplit_job_continuous_and_grid(kind):
    """shard=(rank, world) for the single-group env kinds: rank r
    owns the global ids [r N, (r + 1) N) -- two shards == the unsplit batch."""
    import torch
    from tests.test_cuda_discrete import make_env
    if kind == "continuous":
        cfg = dict(seed=0, state_space_type="continuous", state_space_dim=2,
                   transition_dynamics_order=1, inertia=1.0, time_unit=1.0,
                   target_point=[0.0, 0.0], state_space_max=5.0, action_space_max=1.0,
                   transition_noise=0.1, reward_noise=0.5)
        acts = torch.rand((20, 600, 2), device="cuda") * 2 - 1
    else:
        cfg = dict(seed=0, state_space_type="grid", grid_shape=(8, 8), delay=0,
                   sequence_length=1, reward_function="move_to_a_point",
                   target_point=[5, 5], make_denser=True, transition_noise=0.2,
                   reward_noise=0.5)
        acts = torch.zeros((20, 600, 2), dtype=torch.int64, device="cuda")
        acts[..., 1] = torch.randint(-1, 2, (20, 600), device="cuda")
    whole = make_env(600, autoreset=True, horizon=7, **cfg)
    ref = whole.rollout(20, actions=acts, want_final_obs=False)
    for r in range(2):
        part = make_env(300, autoreset=True, horizon=7, shard=(r, 2), **cfg)
        out = part.rollout(20, actions=acts[:, 300 * r:300 * (r + 1)].contiguous(),
                           want_final_obs=False)
        for k in ref:
            assert torch.equal(out[k], ref[k][:, 300 * r:300 * (r + 1)]), (k, r)


def test_strict_false_treats_foreign_action_dtypes_like_the_reference():
    """Box.contains needs can_cast(action dtype, dtype_s): float64 actions for
    a float32 env freeze the state (rl_toy_env.py:1640, :1671-1679), float
    grid actions are no-ops (:1730-1733).  Default (strict) raises."""
    import torch
    from tests.test_cuda_discrete import make_env
    cfg = dict(seed=0, state_space_type="continuous", state_space_dim=2,
               transition_dynamics_order=2, inertia=1.0, time_unit=0.5,
               target_point=[0.0, 0.0], state_space_max=5.0, action_space_max=1.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ScalarRLToyEnv(**dict(cfg))
    env = make_env(1, noise="numpy", strict=False, **dict(cfg))
    assert np.array_equal(env.curr_obs[0].cpu().numpy(), ref.curr_obs)
    rng = np.random.default_rng(2)
    for t in range(12):
        a = rng.uniform(-1, 1, size=2)
        a = a.astype(np.float32) if t % 3 else a      # every third: float64
        o1, r1, d1, _, _ = ref.step(a)
        o2, r2, d2, _, _ = env.step(torch.as_tensor(a)[None])
        assert np.array_equal(o2[0].cpu().numpy(), o1), t
        assert float(r2[0]) == float(np.float32(r1)), t
    strict = make_env(1, **dict(cfg))
    with pytest.raises(TypeError, match="strict=False"):
        strict.step(torch.zeros((1, 2), dtype=torch.float64))
    strict.step(torch.zeros((1, 2), dtype=torch.float16))   # castable: accepted
    g = make_env(4, strict=False, seed=0, state_space_type="grid", grid_shape=(8, 8),
                 delay=0, sequence_length=1, reward_function="move_to_a_point",
                 target_point=[5, 5], make_denser=True)
    before = g.get_augmented_state()["curr_state"].clone()
    obs, r, term, trunc, _ = g.step(torch.ones((4, 2), dtype=torch.float32))
    # a no-op: cells only clamp back into the grid (reset draws one past it)
    assert torch.equal(obs, before.clamp(max=7))
