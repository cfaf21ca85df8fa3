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
