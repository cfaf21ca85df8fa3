"""GPU parity tests of the discrete step path: CUDA kernels (through the
C ABI) against the reference's golden vectors, the scalar oracle and the
batched oracle.  Bit-exact for states / flags / fp64 rewards whenever the
noise is off or replayed; native Philox noise: states bit-exact against the
oracle's Philox restatement, rewards to libm accuracy, and distribution
tests (chi-squared / KS) with the thresholds written below."""
import warnings

import numpy as np
import pytest
import torch

from oracle.scalar_env import ScalarRLToyEnv
from oracle.vector_oracle import VectorDiscreteOracle
from tests import golden_util as gu
from tests.golden.cases import CASES
from tests.test_vector_oracle import CASES_D, replay_golden_through

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


@pytest.mark.parametrize("name", CASES_D)
def test_cuda_replays_reference_golden(name):
    """Recorded reference draws replayed through the kernels: bit-exact."""
    g = gu.load(name)
    K = g["done"].shape[0]
    env = make_env(K, noise="replay", **gu.case_config(name))
    assert np.array_equal(env.transition_matrix, g["P"])

    def vec_reset(mask, reset_u):
        obs, _ = env.reset(options={"mask": mask, "reset_u": reset_u})
        return obs.cpu().numpy()

    def vec_step(a, u, n):
        obs, r, term, trunc, _ = env.step(
            a, replay=dict(transition_u=np.nan_to_num(u),
                           reward_noise=np.nan_to_num(n)))
        assert not trunc.any()
        return obs.cpu().numpy(), r.cpu().numpy(), term.cpu().numpy()

    replay_golden_through(vec_reset, vec_step, g)


@pytest.mark.parametrize("name", ["c1_seq1", "c2_seq3_del2_noise",
                                  "rdist_scale_shift", "custom_8x5",
                                  "diam3_seq4"])
def test_same_seed_drop_in_numpy_streams(name):
    """noise='numpy': one env, same config and seed as the (oracle of the)
    reference => the same trajectory, constructor included."""
    cfg = gu.case_config(name)
    ref = scalar_oracle(gu.case_config(name))
    env = make_env(1, noise="numpy", **cfg)
    assert int(env.curr_obs[0]) == int(ref.curr_obs)
    rng = np.random.default_rng(9)
    for t in range(150):
        a = int(rng.integers(ref.action_space_size))
        o1, r1, d1, _, _ = ref.step(a)
        o2, r2, d2, tr2, _ = env.step([a])
        assert int(o2[0]) == int(o1) and float(r2[0]) == float(r1), t
        assert bool(d2[0]) == d1
        if d1 or t % 20 == 19:
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert int(o2[0]) == int(o1)


@pytest.mark.parametrize("name,N,T,autoreset,horizon", [
    ("c2_every1", 1500, 40, True, 9),
    ("c2_seq3_del2_noise", 700, 33, True, 0),
    ("big50", 300, 30, True, 11),
    ("diam3_seq4", 257, 25, False, 0),
    ("custom_8x5", 300, 20, True, 5),
    ("notmax_diam2", 300, 40, True, 13),
])
def test_philox_rollout_matches_oracle(name, N, T, autoreset, horizon):
    cfg = gu.case_config(name)
    ora = VectorDiscreteOracle(scalar_oracle(gu.case_config(name)), N,
                               autoreset=autoreset, horizon=horizon, seed=77,
                               env_id_offset=1000)
    env = make_env(N, autoreset=autoreset, horizon=horizon, philox_seed=77,
                   env_id_offset=1000, **cfg)
    ora.reset()  # mirrors the reset at the end of the env constructor
    assert np.array_equal(env._cur.cpu().numpy(), ora.cur)
    for part, acts in ((T, None), (7, "given")):
        actions = None
        if acts:
            actions = np.random.default_rng(3).integers(
                0, ora.A, size=(part, N))
        want = ora.rollout(part, actions=actions)
        got = env.rollout(part, actions=actions)
        for k in ("obs", "final_obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12)
    st = env.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert st[k][0] == ora.stats[k], k
    for k in ("reward", "abs_reward_noise"):
        np.testing.assert_allclose(st[k][0], ora.stats[k], rtol=1e-9)


@pytest.mark.parametrize("name", ["c2_every1", "big50"])
def test_fast_normal_matches_oracle_within_1e5(name):
    """normal_precision='fast' (SFU Box-Muller): states exact, rewards within
    1e-5 absolute of the oracle's fp32 restatement."""
    cfg = gu.case_config(name)
    N, T = 1000, 24
    ora = VectorDiscreteOracle(scalar_oracle(gu.case_config(name)), N,
                               autoreset=True, horizon=10, seed=5,
                               fast_normal=True)
    env = make_env(N, autoreset=True, horizon=10, philox_seed=5,
                   normal_precision="fast", **cfg)
    ora.reset()
    want, got = ora.rollout(T), env.rollout(T)
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=0, atol=1e-5 * max(1.0, cfg.get("reward_noise", 1.0)))


@pytest.mark.parametrize("name", ["c2_every1", "c1_seq1", "big50", "custom_8x5",
                                  "notmax_diam2"])
def test_jit_and_aot_kernels_agree(name):
    """The NVRTC-specialised kernel and the ahead-of-time kernel are the same
    source: identical outputs, bit for bit."""
    outs = []
    for jit in (True, False):
        env = make_env(777, autoreset=True, horizon=9, philox_seed=11,
                       **gu.case_config(name))
        env.set_jit(jit)
        acts = torch.randint(0, env.tables.n_actions, (50, 777),
                             dtype=torch.int32, device="cuda",
                             generator=torch.Generator("cuda").manual_seed(1))
        outs.append(env.rollout(50, actions=acts, want_final_obs=False))
        assert env.jit_last_used == jit, env.jit_log
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_rollout_equals_repeated_steps():
    cfg = gu.case_config("c2_every1")
    a = make_env(513, autoreset=True, horizon=10, **cfg)
    b = make_env(513, autoreset=True, horizon=10, **gu.case_config("c2_every1"))
    ra = a.rollout(19)
    for t in range(19):
        acts = None
        rb = b.rollout(1)
        for k in ra:
            assert torch.equal(ra[k][t], rb[k][0]), (k, t)


def test_sharding_invariance():
    """Global env ids key the noise: 2 shards == 1 big batch."""
    cfg = lambda: gu.case_config("c2_every1")  # noqa: E731
    whole = make_env(600, autoreset=True, horizon=12, **cfg())
    lo = make_env(300, autoreset=True, horizon=12, env_id_offset=0, **cfg())
    hi = make_env(300, autoreset=True, horizon=12, env_id_offset=300, **cfg())
    rw, rl, rh = whole.rollout(30), lo.rollout(30), hi.rollout(30)
    for k in rw:
        assert torch.equal(rw[k], torch.cat([rl[k], rh[k]], dim=1)), k


def test_augmented_state_and_masked_reset():
    cfg = gu.case_config("c2_every1")
    ref = scalar_oracle(gu.case_config("c2_every1"))
    env = make_env(3, noise="numpy", **cfg)
    for t in range(9):
        a = t % 8
        ref.step(a)
        env.step([a, a, a])
        aug = env.get_augmented_state()["augmented_state"][0].cpu().numpy()
        want = np.array(ref.augmented_state, dtype=np.float64)
        assert np.array_equal(np.isnan(aug), np.isnan(want)), (t, aug, want)
        assert np.array_equal(np.nan_to_num(aug), np.nan_to_num(want))
    before = env._cur.clone()
    env.reset(options={"mask": [False, True, False]})
    assert int(env._t[0]) == 9 and int(env._t[1]) == 0
    assert int(env._cur[0]) == int(before[0])


def test_transition_noise_distribution_chi2():
    """chi-squared of the noisy next-state histogram against the reference's
    probabilities (1-p at P[s,a], p/(S-1) elsewhere); 7 dof, reject at
    p < 1e-4 (statistic > 29.9)."""
    cfg = dict(gu.case_config("c1_seq1"), transition_noise=0.25,
               terminal_state_density=0.0)
    N = 200000
    env = make_env(N, philox_seed=123, **cfg)
    s0 = 3
    env.reset(options={"init_state": np.full(N, s0)})
    out = env.rollout(1, actions=np.full((1, N), 2))
    nxt = int(env.transition_matrix[s0, 2])
    counts = np.bincount(out["final_obs"][0].cpu().numpy(), minlength=8)
    probs = np.full(8, 0.25 / 7)
    probs[nxt] = 0.75
    chi2 = float(((counts - N * probs) ** 2 / (N * probs)).sum())
    assert chi2 < 29.9, (chi2, counts)
    assert env.episode_stats()["noisy_transitions"][0] == N - counts[nxt]


@pytest.mark.parametrize("precision", ["fp64", "fast"])
def test_reward_noise_distribution_ks(precision):
    """Kolmogorov-Smirnov of the reward noise against N(0, sigma); reject at
    p < 1e-4 (D * sqrt(n) > 2.23).  Also mean/variance within 5 sigma."""
    from scipy import special
    sigma = 0.25
    # reward_every_n_steps larger than the run gates every sequence reward
    # to 0, so the returned reward is exactly the noise draw
    cfg = dict(gu.case_config("c1_seq1"), reward_noise=sigma,
               reward_every_n_steps=10**6, terminal_state_density=0.0)
    N = 200000
    env = make_env(N, philox_seed=321, normal_precision=precision, **cfg)
    out = env.rollout(5, want_final_obs=False)
    r = out["reward"].cpu().numpy().ravel()
    z = r
    zs = np.sort(z) / sigma
    cdf = special.ndtr(zs)
    n = zs.size
    d = max(np.max(cdf - np.arange(n) / n), np.max(np.arange(1, n + 1) / n - cdf))
    assert d * np.sqrt(n) < 2.23, d
    assert abs(z.mean()) < 5 * sigma / np.sqrt(z.size)
    assert abs(z.var() - sigma ** 2) < 5 * sigma ** 2 * np.sqrt(2 / z.size)


def test_init_state_distribution_and_full_size_properties():
    """BASELINE config #2 at full size (65 536 envs x 1000 steps): properties
    that need no oracle run -- states in range, terminated <=> terminal
    state, truncation exactly at the horizon, auto-reset lands on
    non-terminal states uniformly (chi-squared, 5 dof, < 25.7 at p=1e-4),
    delayed reward pattern: nothing paid before t > delay."""
    cfg = gu.case_config("c2_every1")
    N, T, H = 65536, 1000, 100
    env = make_env(N, autoreset=True, horizon=H, **cfg)
    out = env.rollout(T)
    obs, fin = out["obs"], out["final_obs"]
    term, trunc = out["terminated"], out["truncated"]
    assert int(fin.min()) >= 0 and int(fin.max()) < 8
    mask = torch.zeros(8, dtype=torch.bool, device=fin.device)
    mask[torch.as_tensor(env.tables.terminal_states, device=fin.device)] = True
    assert torch.equal(term, mask[fin])
    reset = term | trunc
    assert torch.equal(obs[~reset], fin[~reset])
    assert not mask[obs[reset]].any()
    counts = torch.bincount(obs[reset], minlength=8).cpu().numpy()[:6]
    exp = counts.sum() / 6
    assert ((counts - exp) ** 2 / exp).sum() < 25.7
    st = env.episode_stats()
    assert st["transitions"][0] == N * T
    assert st["episodes"][0] == int(reset.sum())
    assert st["terminated"][0] == int(term.sum())


@pytest.mark.parametrize("name", ["c2_every1", "c4_img_all"])
def test_graphed_step_equals_eager_step(name):
    """A CUDA-graph replay of step() (+ renderer) == eager step() calls, also
    when eager calls are mixed in between replays."""
    cfg = gu.case_config(name)
    a = make_env(300, autoreset=True, horizon=6, philox_seed=4, **cfg)
    b = make_env(300, autoreset=True, horizon=6, philox_seed=4,
                 **gu.case_config(name))
    fn = b.make_graphed_step()
    assert torch.equal(a._cur, b._cur)
    rng = np.random.default_rng(0)
    for t in range(20):
        acts = torch.as_tensor(rng.integers(0, 8, size=300), dtype=torch.int32,
                               device="cuda")
        ra = a.step(acts)
        rb = fn(acts) if t % 5 != 4 else b.step(acts)
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y), t


def test_set_augmented_state_round_trip_and_bare_state():
    """get_augmented_state() -> set_augmented_state() on a fresh env continues
    the same trajectory (delay 0, so no hidden FIFO contents); a bare state
    behaves like a fresh episode in that state (:2168-2215)."""
    cfg = dict(gu.case_config("c2_every1"), delay=0)
    a = make_env(64, philox_seed=3, **dict(cfg))
    b = make_env(64, philox_seed=3, **dict(cfg))
    acts = torch.randint(0, 8, (9, 64), dtype=torch.int32, device="cuda")
    a.rollout(5, actions=acts[:5])
    b._step_index = a._step_index            # same Philox clock
    b.set_augmented_state(a.get_augmented_state())
    ra, rb = a.rollout(4, actions=acts[5:]), b.rollout(4, actions=acts[5:])
    for k in ra:
        assert torch.equal(ra[k], rb[k]), k
    ref = scalar_oracle(dict(cfg))
    c = make_env(1, noise="numpy", **dict(cfg))
    ref.augmented_state = [np.nan] * (len(ref.augmented_state) - 1) + [2]
    ref.curr_state = 2
    c.set_augmented_state(torch.tensor([2]))
    for t, act in enumerate([1, 5, 3, 3, 0, 6]):
        o1, r1, d1, _, _ = ref.step(act)
        o2, r2, d2, _, _ = c.step([act])
        assert int(o2[0]) == int(o1) and float(r2[0]) == float(r1), t


# ---------------------------------------------------------------------------
# the exact bench.py call (headline kernel) against the oracle
# ---------------------------------------------------------------------------
_BENCH_CFG = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
                  state_space_size=8, action_space_size=8, sequence_length=3,
                  delay=2, transition_noise=0.1, reward_noise=0.25,
                  reward_density=0.25, terminal_state_density=0.25,
                  generate_random_mdp=True, reward_every_n_steps=True)
_bench_oracle_cache = {}


def _bench_oracle(normal, prologue, T, N):
    key = (normal, prologue, T, N)
    if key not in _bench_oracle_cache:
        ora = VectorDiscreteOracle(
            scalar_oracle(dict(_BENCH_CFG)), N, autoreset=True, horizon=100,
            seed=0, env_id_offset=5 * N, fast_normal=normal == "fast",
            normal="boxmuller" if normal == "boxmuller" else "ziggurat")
        ora.reset()
        acts = np.random.default_rng(0xC0FFEE).integers(0, 8, size=(prologue + T, N))
        if prologue:
            ora.rollout(prologue, actions=acts[:prologue])
        _bench_oracle_cache[key] = (acts, ora.rollout(T, actions=acts[prologue:]),
                                    dict(ora.stats))
    return _bench_oracle_cache[key]


@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("normal", ["fp64", "boxmuller", "fast"])
@pytest.mark.parametrize("prologue,T", [(0, 107), (3, 110)])
def test_bench_signature_matches_oracle(jit, normal, prologue, T):
    """BASELINE config #2 through the call bench.py times -- int32 actions
    given, `out` given, no final_obs: the FAST / standard-signature kernel
    with the chunk-8 action-prefetch loop -- against the batched oracle with
    noise ON.  (0, 107): 13 full chunks + 3 single steps; (3, 110): 1 peeled
    step, 13 full chunks, a half chunk and a single step.  States and flags
    bit-exact; rewards 1e-12 (fp64 normals; the ziggurat's 98.5 % fast path is
    exact, its tail goes through log1p) or 1e-5 (fp32 SFU normals)."""
    N = 4096
    acts, want, stats = _bench_oracle(normal, prologue, T, N)
    env = make_env(N, autoreset=True, horizon=100, env_id_offset=5 * N,
                   normal_precision=normal, **dict(_BENCH_CFG))
    env.set_jit(jit)
    a = torch.as_tensor(acts, dtype=torch.int32, device="cuda")
    if prologue:
        env.rollout(prologue, actions=a[:prologue])
    out = {"obs": torch.empty((T, N), dtype=torch.int64, device="cuda"),
           "reward": torch.empty((T, N), dtype=torch.float64, device="cuda"),
           "terminated": torch.empty((T, N), dtype=torch.bool, device="cuda"),
           "truncated": torch.empty((T, N), dtype=torch.bool, device="cuda")}
    got = env.rollout(T, actions=a[prologue:].contiguous(), out=out)
    assert env.jit_last_used == jit, env.jit_log
    for k in ("obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    tol = 1e-5 if normal == "fast" else 1e-12
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=0 if normal == "fast" else tol, atol=tol)
    assert want["terminated"].sum() > 1000
    st = env.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert st[k][0] == stats[k], k


@pytest.mark.parametrize("prologue,T", [(0, 300), (3, 427)])
def test_long_launch_standard_signature_matches_oracle(prologue, T):
    """The bench signature over MANY staged ziggurat windows (T = 300 / 427:
    9 / 13 windows of 32 steps, peeled prologue step, partial last window),
    run-time specialised and ahead-of-time kernel, against the oracle: states
    / flags exact, rewards 1e-12, and JIT == AOT bit for bit.  (Written for the
    pipelined-staging experiment, tools/experiments/: any restructuring of
    where the normals are drawn must keep the Philox words per (env, step).)"""
    N = 2048
    acts, want, stats = _bench_oracle("fp64", prologue, T, N)
    a = torch.as_tensor(acts, dtype=torch.int32, device="cuda")
    outs = []
    for jit in (True, False):
        env = make_env(N, autoreset=True, horizon=100, env_id_offset=5 * N,
                       normal_precision="fp64", **dict(_BENCH_CFG))
        env.set_jit(jit)
        if prologue:
            env.rollout(prologue, actions=a[:prologue])
        got = env.rollout(T, actions=a[prologue:].contiguous(), want_final_obs=False)
        assert env.jit_last_used == jit, env.jit_log
        for k in ("obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (jit, k)
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12)
        st = env.episode_stats()
        for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
            assert st[k][0] == stats[k], (jit, k)
        outs.append(got["reward"])
    assert torch.equal(outs[0], outs[1])  # NVRTC-specialised == ahead-of-time


def test_ziggurat_reward_noise_is_standard_normal():
    """KS test of the native fp64 (ziggurat) reward noise, 2M draws incl. the
    wedge / tail path: (reward - noise-free reward) / sigma ~ N(0, 1)."""
    from scipy import stats
    N, T = 8192, 256
    cfg = dict(_BENCH_CFG, transition_noise=0)
    a = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda",
                      generator=torch.Generator("cuda").manual_seed(2))
    noisy = make_env(N, autoreset=True, horizon=50, philox_seed=4, **cfg)
    clean = make_env(N, autoreset=True, horizon=50, philox_seed=4,
                     **dict(cfg, reward_noise=0))
    r1 = noisy.rollout(T, actions=a, want_final_obs=False)
    r0 = clean.rollout(T, actions=a, want_final_obs=False)
    assert torch.equal(r1["obs"], r0["obs"])
    z = ((r1["reward"] - r0["reward"]) / 0.25).flatten().cpu().numpy()
    assert stats.kstest(z, "norm").pvalue > 1e-4
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1) < 3e-3
    assert (np.abs(z) > 3.6541528853610088).sum() > 300   # tail draws present


def test_oracle_adopts_device_state_mid_run():
    """VectorDiscreteOracle.load_state (what bench.py's in-run parity check
    uses): adopt the CUDA path's SoA state after 57 steps for a scattered
    sample of envs, then both continue for 43 steps: identical."""
    N = 2048
    env = make_env(N, autoreset=True, horizon=100, env_id_offset=3 * N,
                   **dict(_BENCH_CFG))
    a = torch.randint(0, 8, (100, N), dtype=torch.int32, device="cuda",
                      generator=torch.Generator("cuda").manual_seed(5))
    env.rollout(57, actions=a[:57], want_final_obs=False)
    idx = np.array([0, 1, 63, 64, 777, 1024, 2047])
    ora = VectorDiscreteOracle(scalar_oracle(dict(_BENCH_CFG)), len(idx),
                               autoreset=True, horizon=100, seed=0,
                               gids=3 * N + idx)
    ora.load_state(env._cur.cpu().numpy()[idx], env._t.cpu().numpy()[idx],
                   env._episode.cpu().numpy()[idx], env._key.cpu().numpy()[idx],
                   env._ring.cpu().numpy()[:, idx], env._step_index, 3)
    got = env.rollout(43, actions=a[57:].contiguous(), want_final_obs=False)
    want = ora.rollout(43, actions=a[57:].cpu().numpy()[:, idx])
    for k in ("obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy()[:, idx], want[k]), k
    np.testing.assert_allclose(got["reward"].cpu().numpy()[:, idx], want["reward"],
                               rtol=1e-12, atol=1e-12)
    assert (want["reward"] != 0).sum() > 100


@pytest.mark.parametrize("dtype_o", [np.uint8, np.int32, np.int16])
def test_dtype_o_written_in_kernel(dtype_o):
    """The reference's dtype_o (rl_toy_env.py:571, :611-614): rollout() and
    step() return observations in it -- uint8 / int32 straight from the kernel
    (JIT and AOT), other dtypes through a cast -- equal to the int64 ones."""
    cfg = dict(_BENCH_CFG)
    N, T = 1000, 30
    a = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda",
                      generator=torch.Generator("cuda").manual_seed(3))
    ref = make_env(N, autoreset=True, horizon=13, **cfg).rollout(T, actions=a)
    want_t = getattr(torch, np.dtype(dtype_o).name)
    for jit in (True, False):
        env = make_env(N, autoreset=True, horizon=13, dtype_o=dtype_o, **cfg)
        env.set_jit(jit)
        if dtype_o is np.int16:
            got = env.rollout(T, actions=a)        # kernel writes int64 ...
            assert got["obs"].dtype == torch.int64
        else:
            got = env.rollout(T, actions=a)
            assert got["obs"].dtype == want_t and got["final_obs"].dtype == want_t
        for k in ("obs", "final_obs"):
            assert torch.equal(got[k].to(torch.int64), ref[k]), k
        assert torch.equal(got["reward"], ref["reward"])
        env2 = make_env(N, autoreset=True, horizon=13, dtype_o=dtype_o, **cfg)
        env2.set_jit(jit)
        for t in range(5):
            obs, r, term, trunc, info = env2.step(a[t])
            assert obs.dtype == want_t                 # ... step() casts it
            assert torch.equal(obs.to(torch.int64), ref["obs"][t])
            assert torch.equal(info["final_obs"].to(torch.int64), ref["final_obs"][t])
            assert torch.equal(r, ref["reward"][t])


def test_step_buffers_rotate_and_match_unbuffered_step():
    """step() writes into `step_buffers` rotating pre-marshalled output sets:
    same values as the allocate-per-call path, and a returned tensor stays
    valid for step_buffers - 1 further calls."""
    cfg = dict(_BENCH_CFG)
    N = 513
    a = torch.randint(0, 8, (12, N), dtype=torch.int32, device="cuda",
                      generator=torch.Generator("cuda").manual_seed(8))
    fast = make_env(N, autoreset=True, horizon=7, step_buffers=3, **cfg)
    slow = make_env(N, autoreset=True, horizon=7, step_buffers=0, **cfg)
    kept = []
    for t in range(12):
        f, s = fast.step(a[t]), slow.step(a[t])
        for x, y in zip(f[:4], s[:4]):
            assert torch.equal(x, y)
        assert torch.equal(f[4]["final_obs"], s[4]["final_obs"])
        kept.append((f[0], s[0].clone()))
        if t >= 2:  # the tensors of two calls ago are still intact
            assert torch.equal(kept[t - 2][0], kept[t - 2][1])
    assert fast.step(a[0].to(torch.int64))[0].shape == (N,)   # converted on the fly
