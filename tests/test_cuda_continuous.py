"""GPU parity tests of the continuous move_to_a_point path.

Contract (BASELINE.json north_star): states within 1e-5 relative in fp32 and
1e-12 in fp64.  The kernel follows the reference's dtype path exactly, so the
tests demand bit-exact states / derivatives / rewards wherever numpy's result
does not depend on the host BLAS (see is_exact_case), and the contract's
tolerance otherwise."""
import warnings

import numpy as np
import pytest
import torch

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_continuous_oracle import VectorContinuousOracle
from tests import golden_util as gu
from tests.golden.cases import CASES
from tests.test_vector_continuous_oracle import (CASES_C, is_exact_case,
                                                 replay_continuous_golden)

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


@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("name", CASES_C)
def test_cuda_replays_reference_golden(name, jit):
    g = gu.load(name)
    K = g["done"].shape[0]
    env = make_env(K, noise="replay", **gu.case_config(name))
    env.set_jit(jit)

    def vec_reset(mask, init):
        env.reset(options={"mask": mask, "init_state": init})
        return env.get_augmented_state()["curr_state"].cpu().numpy()

    def vec_step(a, sn, rn):
        rep = {"reward_noise": rn}
        if sn is not None:
            rep["state_noise"] = sn
        obs, r, term, trunc, info = env.step(torch.as_tensor(a), replay=rep)
        assert env.jit_last_used == jit, env.jit_log
        derivs = env.get_augmented_state()["state_derivatives"].cpu().numpy()
        return (info["state"].cpu().numpy(), r.cpu().numpy(),
                term.cpu().numpy(), derivs)

    # move_along_a_line: the fitted direction is fp32-accurate, not LAPACK's bits
    line = CASES[name]["config"].get("reward_function") == "move_along_a_line"
    replay_continuous_golden(vec_reset, vec_step, g,
                             exact=is_exact_case(name) and not line)


@pytest.mark.parametrize("name", ["c3_order2", "cont_noise_delay",
                                  "cont_term_boxes", "cont_unbounded"])
def test_same_seed_drop_in_numpy_streams(name):
    """noise='numpy': same config + seed => the reference's trajectory,
    constructor (first reset) included."""
    cfg = gu.case_config(name)
    ref = scalar_oracle(gu.case_config(name))
    env = make_env(1, noise="numpy", **cfg)
    assert np.array_equal(env.curr_obs[0].cpu().numpy(), ref.curr_obs)
    rng = np.random.default_rng(5)
    D = cfg["state_space_dim"]
    amax = cfg.get("action_space_max", 1.0)
    exact = is_exact_case(name)
    for t in range(120):
        a = rng.uniform(-1.05 * amax, 1.05 * amax, size=D).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o1, r1, d1, _, _ = ref.step(a.copy())
        o2, r2, d2, _, _ = env.step(torch.as_tensor(a)[None])
        o2, r2 = o2[0].cpu().numpy(), float(r2[0])
        if exact:
            assert np.array_equal(o2, o1), t
            assert r2 == float(np.float32(r1)), (t, r2, r1)
        else:
            np.testing.assert_allclose(o2, o1, rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(r2, float(r1), rtol=1e-5, atol=1e-6)
        assert bool(d2[0]) == d1
        if d1 or t % 25 == 24:
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert np.array_equal(o2[0].cpu().numpy(), o1)


@pytest.mark.parametrize("name,N,T,autoreset,horizon", [
    ("c3_order2", 3000, 50, True, 20),
    ("cont_noise_delay", 2000, 40, True, 15),
    ("cont_order3", 1000, 40, False, 0),
    ("cont_term_boxes", 2000, 60, True, 25),
    ("cont_sparse", 1500, 40, True, 0),
    ("cont_unbounded", 1000, 30, True, 10),
    ("cont_line_seq10", 600, 45, True, 25),       # move_along_a_line, 4 dims
    ("cont_line_seq3_delay", 600, 40, True, 15),  # 3 of 6 dims, delay, noise
    ("cont_line_2d", 800, 40, True, 0),           # closed-form 2-D direction
    ("cont_inertia_list", 500, 30, True, 10),
])
def test_philox_rollout_matches_oracle(name, N, T, autoreset, horizon):
    """Native Philox noise / reset sampling vs the oracle's restatement of the
    same streams: states 1e-5 relative (contract), typically bit-exact."""
    cfg = gu.case_config(name)
    ora = VectorContinuousOracle(scalar_oracle(gu.case_config(name)), N,
                                 autoreset=autoreset, horizon=horizon, seed=9,
                                 env_id_offset=50)
    env = make_env(N, autoreset=autoreset, horizon=horizon, philox_seed=9,
                   env_id_offset=50, **cfg)
    ora.reset()  # mirrors the reset at the end of the env constructor
    np.testing.assert_allclose(env.curr_obs.cpu().numpy(), ora.em, rtol=1e-6)
    D = cfg["state_space_dim"]
    amax = cfg.get("action_space_max", 1.0)
    acts = np.random.default_rng(2).uniform(
        -1.05 * amax, 1.05 * amax, size=(T, N, D)).astype(np.float32)
    want = ora.rollout(T, acts)
    got = env.rollout(T, torch.as_tensor(acts))
    for k in ("terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    for k in ("obs", "final_obs"):
        np.testing.assert_allclose(got[k].cpu().numpy(), want[k], rtol=1e-5,
                                   atol=1e-6)
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("jit", [True, False])
def test_ziggurat_noise_matches_oracle(jit):
    """normal_precision='ziggurat': numpy's Generator.normal algorithm on
    Philox words for the transition and the reward noise of the continuous
    kernels, against the oracle's restatement of the same streams; its draws
    differ from the default's (fp64 Box-Muller in this kernel)."""
    cfg = gu.case_config("cont_noise_delay")
    N, T = 1500, 30
    ora = VectorContinuousOracle(scalar_oracle(gu.case_config("cont_noise_delay")), N,
                                 autoreset=True, horizon=15, seed=9,
                                 env_id_offset=50, normal="ziggurat")
    env = make_env(N, autoreset=True, horizon=15, philox_seed=9, env_id_offset=50,
                   normal_precision="ziggurat", **cfg)
    env.set_jit(jit)
    ora.reset()
    D = cfg["state_space_dim"]
    amax = cfg.get("action_space_max", 1.0)
    acts = np.random.default_rng(2).uniform(-amax, amax, size=(T, N, D)).astype(np.float32)
    want = ora.rollout(T, acts)
    got = env.rollout(T, torch.as_tensor(acts))
    assert env.jit_last_used == jit
    for k in ("terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(got["obs"].cpu().numpy(), want["obs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-4, atol=1e-5)
    dflt = make_env(N, autoreset=True, horizon=15, philox_seed=9, env_id_offset=50, **cfg)
    other = dflt.rollout(T, torch.as_tensor(acts))
    assert not torch.equal(other["obs"], got["obs"])


def test_continuous_ziggurat_noise_is_standard_normal():
    """normal_precision='ziggurat' on the continuous kernels (numpy's
    ziggurat on Philox words): with zero actions the integrated state
    stays where reset() put it (the noise is added to the emitted state only,
    rl_toy_env.py:1683-1700), so obs - initial obs = the noise of that step;
    the per-dimension draws pass a KS test against N(0, sigma) and are
    uncorrelated."""
    from scipy import stats
    cfg = dict(seed=0, state_space_type="continuous", state_space_dim=3,
               transition_dynamics_order=1, inertia=1.0, time_unit=1.0,
               target_point=[0.0, 0.0, 0.0], target_radius=1e-9,
               state_space_max=1e6, action_space_max=1.0, transition_noise=0.5,
               reward_noise=2.0, make_denser=True, reward_scale=1.0, dtype_s=np.float64)
    N, T = 20000, 8
    env = make_env(N, autoreset=False, philox_seed=3, normal_precision="ziggurat", **cfg)
    acts = torch.zeros((T, N, 3), dtype=torch.float64, device="cuda")
    obs0 = env.curr_obs.clone()
    out = env.rollout(T, actions=acts)
    prev = torch.cat([obs0[None], out["obs"][:-1]])
    noise = (out["obs"] - obs0[None]).cpu().numpy().reshape(-1, 3)  # zero action: pure noise
    for d in range(3):
        assert stats.kstest(noise[:, d] / 0.5, "norm").pvalue > 1e-3, d
    assert abs(np.corrcoef(noise.T)[0, 1]) < 0.01
    # dense reward = distance moved towards the target + N(0, 2): the noise part
    d_prev = np.linalg.norm(prev.cpu().numpy(), axis=-1)
    d_new = np.linalg.norm(out["obs"].cpu().numpy(), axis=-1)
    rn = out["reward"].cpu().numpy() - (d_prev - d_new)
    assert stats.kstest(rn.reshape(-1) / 2.0, "norm").pvalue > 1e-3


def test_fast_normal_noise_matches_oracle():
    """normal_precision='fast' (SFU Box-Muller) on the continuous path: the
    oracle restates the same fp32 Box-Muller; 1e-5 contract tolerance (the SFU
    approximations differ from numpy's libm in the last bits, and transition
    noise feeds back into the state)."""
    cfg = gu.case_config("cont_noise_delay")
    N, T = 2000, 40
    ora = VectorContinuousOracle(scalar_oracle(gu.case_config("cont_noise_delay")), N,
                                 autoreset=True, horizon=15, seed=9,
                                 env_id_offset=50, fast_normal=True)
    env = make_env(N, autoreset=True, horizon=15, philox_seed=9, env_id_offset=50,
                   normal_precision="fast", **cfg)
    ora.reset()
    D = cfg["state_space_dim"]
    amax = cfg.get("action_space_max", 1.0)
    acts = np.random.default_rng(2).uniform(-amax, amax, size=(T, N, D)).astype(np.float32)
    want = ora.rollout(T, acts)
    got = env.rollout(T, torch.as_tensor(acts))
    # a 1e-6 difference can flip a terminal test on a boundary: compare where
    # the trajectories agree on the flags, and require that almost everywhere
    same = np.array_equal(got["terminated"].cpu().numpy(), want["terminated"])
    agree = (got["terminated"].cpu().numpy() == want["terminated"]).mean()
    assert same or agree > 0.9999
    ok = np.isclose(got["obs"].cpu().numpy(), want["obs"], rtol=1e-4, atol=1e-4)
    assert ok.mean() > 0.999
    okr = np.isclose(got["reward"].cpu().numpy(), want["reward"], rtol=1e-3, atol=1e-4)
    assert okr.mean() > 0.999


def test_fp64_build_matches_oracle_1e12():
    """dtype_s=float64 (the fp64 verification build): 1e-12 relative."""
    cfg = dict(gu.case_config("c3_order2"), dtype_s=np.float64,
               transition_noise=0.01, reward_noise=0.1)
    N, T = 500, 60
    ora = VectorContinuousOracle(scalar_oracle(dict(cfg)), N, autoreset=True,
                                 horizon=30, seed=4)
    env = make_env(N, autoreset=True, horizon=30, philox_seed=4, **dict(cfg))
    ora.reset()
    acts = np.random.default_rng(8).uniform(-1, 1, size=(T, N, 6))
    want = ora.rollout(T, acts)
    got = env.rollout(T, torch.as_tensor(acts))
    assert got["obs"].dtype == torch.float64
    assert np.array_equal(got["terminated"].cpu().numpy(), want["terminated"])
    np.testing.assert_allclose(got["obs"].cpu().numpy(), want["obs"],
                               rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-12, atol=1e-12)  # the contract's fp64 bar


def test_full_size_config3_properties():
    """BASELINE config #3 at full size (1M envs): size-independent checks --
    in-range constant action moves every env identically relative to its
    start (linearity of the order-2 dynamics), states stay inside the box,
    the dense rewards telescope to the change in distance to the target."""
    cfg = gu.case_config("c3_order2")
    N, T = 1 << 20, 20
    env = make_env(N, **cfg)
    s0 = env.curr_obs.clone()
    a = torch.full((T, N, 6), 0.25, dtype=torch.float32, device="cuda")
    out = env.rollout(T, a, want_final_obs=False)
    obs = out["obs"]
    assert float(obs.abs().max()) <= 10.0
    # order 2, inertia 1, tu 0.5: x(T) - x(0) = a * (T*tu)^2 / 2 while unclipped
    disp = obs[-1] - s0
    inside = (obs.abs().amax(dim=(0, 2)) < 10.0)
    want = 0.25 * (T * 0.5) ** 2 / 2
    assert torch.allclose(disp[inside], torch.full_like(disp[inside], want),
                          rtol=1e-4, atol=1e-4)
    d0 = s0[:, :2].norm(dim=1)
    dT = obs[-1][:, :2].norm(dim=1)
    never_done = ~out["terminated"].any(dim=0)
    tot = out["reward"].sum(dim=0)
    assert torch.allclose(tot[never_done], (d0 - dT)[never_done], atol=1e-3)


def test_graphed_step_equals_eager_step_continuous():
    cfg = gu.case_config("cont_noise_delay")
    a = make_env(200, autoreset=True, horizon=9, philox_seed=4, **cfg)
    b = make_env(200, autoreset=True, horizon=9, philox_seed=4,
                 **gu.case_config("cont_noise_delay"))
    fn = b.make_graphed_step()
    rng = np.random.default_rng(0)
    for t in range(15):
        acts = torch.as_tensor(rng.uniform(-1, 1, size=(200, 6)).astype(np.float32),
                               device="cuda")
        ra, rb = a.step(acts), fn(acts)
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y), t
