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
    grid_numpy_leg(gu.load(name), gu.case_config(name),
                   CASES[name].get("horizon", 12))


def grid_numpy_leg(g, cfg, H):
    """(also run on freshly recorded cases by tests/test_fuzz_reference.py)"""
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
        _check_lane(env, g, k, H, image)


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


# --------------------------------------------------------------------------
# batched oracle (CPU) and the CUDA path (GPU)
# --------------------------------------------------------------------------
from oracle.vector_grid_oracle import VectorGridOracle  # noqa: E402

GRID_PLAIN = [n for n in gu.GRID_CASES
              if not CASES[n]["config"].get("image_representations")]


def make_env(*a, **k):
    from mdp_playground_b200 import VectorRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return VectorRLToyEnv(*a, **k)


def replay_grid_golden(vec_reset, vec_step, g):
    """Every golden lane = one env of the batch; recorded draws replayed,
    resets issued where the reference reset."""
    K, T = g["done"].shape
    cur = vec_reset(None, g["init_state"], None)
    assert np.array_equal(cur, g["init_state"])
    for t in range(T):
        obs, r, done = vec_step(g["actions"][:, t], dict(
            noise_u=np.nan_to_num(g["grid_noise_u"][:, t], nan=1.0),
            noise_action=g["grid_noise_action"][:, t],
            reward_noise=np.nan_to_num(g["reward_noise"][:, t])), t)
        assert np.array_equal(obs, g["state"][:, t]), t
        assert np.array_equal(r, g["reward"][:, t]), (t, r, g["reward"][:, t])
        assert np.array_equal(done, g["done"][:, t]), t
        m = g["reset_after"][:, t]
        if m.any():
            cur = vec_reset(m, np.where(m[:, None], g["reset_state"][:, t],
                                        g["state"][:, t]), t)
            assert np.array_equal(cur[m], g["reset_state"][m, t]), t


@pytest.mark.parametrize("name", GRID_PLAIN)
def test_vector_oracle_replays_reference_golden(name):
    g = gu.load(name)
    vec = VectorGridOracle(scalar_oracle(gu.case_config(name)), g["done"].shape[0])

    def vec_reset(mask, init, t):
        return vec.reset(mask=mask, init_state=init)

    def vec_step(a, rep, t):
        out = vec.rollout(1, a[None], replay={k: v[None] for k, v in rep.items()})
        return out["obs"][0], out["reward"][0], out["terminated"][0]

    replay_grid_golden(vec_reset, vec_step, g)


def test_substitute_action_is_uniform_over_the_other_samples():
    """Closed form of the reference's rejection loop: every GridActionSpace
    sample (dim, value) other than the current action equally likely."""
    from oracle.vector_grid_oracle import substitute_action
    for nd in (2, 4):
        for a in ([0] * nd, [1] + [0] * (nd - 1), [0] * (nd - 1) + [-1]):
            ws = np.linspace(0, 2 ** 32 - 1, 6 * nd * 50).astype(np.uint64)
            counts = {}
            for w in ws:
                b = tuple(substitute_action(w, a))
                assert list(b) != a and sum(abs(v) for v in b) <= 1
                counts[b] = counts.get(b, 0) + 1
            moves = {k: v for k, v in counts.items() if any(k)}
            assert len(moves) == 2 * nd - (1 if any(a) else 0)
            assert max(moves.values()) - min(moves.values()) <= 2
            if any(a):  # the no-op stands for nd of the 3*nd - 1 samples
                assert abs(counts[(0,) * nd] - nd * np.mean(list(moves.values()))) <= 2 * nd


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.GRID_CASES)
def test_cuda_replays_reference_golden(name):
    g = gu.load(name)
    K = g["done"].shape[0]
    image = "obs_image" in g
    env = make_env(K, noise="replay", **gu.case_config(name))

    def vec_reset(mask, init, t):
        obs, _ = env.reset(options={"mask": mask, "init_state": init})
        if image:
            want = g["init_image"] if t is None else g["reset_image"][:, t]
            sel = slice(None) if mask is None else mask
            assert np.array_equal(obs.cpu().numpy()[sel], want[sel])
        return env.get_augmented_state()["curr_state"].cpu().numpy()

    def vec_step(a, rep, t):
        obs, r, term, trunc, info = env.step(a, replay=rep)
        assert not trunc.any()
        if image:
            assert np.array_equal(obs.cpu().numpy(), g["obs_image"][:, t]), t
        return info["state"].cpu().numpy(), r.cpu().numpy(), term.cpu().numpy()

    replay_grid_golden(vec_reset, vec_step, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.GRID_CASES)
def test_same_seed_drop_in_numpy_streams(name):
    """noise='numpy': one env, same config and seed as the (oracle of the)
    reference => the same trajectory, constructor included."""
    cfg = gu.case_config(name)
    ref = scalar_oracle(gu.case_config(name))
    env = make_env(1, noise="numpy", **cfg)
    nd = len(ref.grid_shape)
    assert np.array_equal(env.curr_obs[0].cpu().numpy(), ref.curr_obs)
    rng = np.random.default_rng(9)
    for t in range(40 if cfg.get("image_representations") else 200):
        a = [0] * nd
        if rng.integers(8) < 7:
            d = int(rng.integers(nd))
            a[d] = int(np.sign(ref.target_point[d % 2] - ref.curr_state[d])) \
                if rng.integers(2) else int(rng.integers(-1, 2))
        else:
            a = [int(v) for v in rng.integers(-1, 3, size=nd)]
        o1, r1, d1, _, _ = ref.step(list(a))
        o2, r2, d2, tr2, info = env.step(np.array([a]))
        assert np.array_equal(o2[0].cpu().numpy(), o1), t
        assert float(r2[0]) == float(r1) and bool(d2[0]) == d1, t
        assert np.array_equal(info["state"][0].cpu().numpy(), ref.curr_state)
        if t % 25 == 24 or (d1 and t % 2):
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert np.array_equal(o2[0].cpu().numpy(), o1)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,T,autoreset,horizon", [
    ("grid_dense_term", 1000, 40, True, 9),
    ("grid_sparse_noise", 777, 33, True, 0),
    ("grid_irr_5x9", 300, 50, True, 11),
    ("grid_sparse_noise", 257, 25, False, 0),
])
def test_philox_rollout_matches_oracle(name, N, T, autoreset, horizon):
    cfg = gu.case_config(name)
    ora = VectorGridOracle(scalar_oracle(gu.case_config(name)), N,
                           autoreset=autoreset, horizon=horizon, seed=77,
                           env_id_offset=1000)
    env = make_env(N, autoreset=autoreset, horizon=horizon, philox_seed=77,
                   env_id_offset=1000, **cfg)
    first = ora.reset()
    assert np.array_equal(env.get_augmented_state()["curr_state"].cpu().numpy(), first)
    shape = np.array(ora.shape)
    assert (first == shape).any() and (first <= shape).all()  # one-past cells occur
    rng = np.random.default_rng(3)
    for part in (T, 7, 1):
        acts = np.zeros((part, N, ora.nd), dtype=np.int64)
        d = rng.integers(ora.nd, size=(part, N))
        np.put_along_axis(acts, d[..., None], rng.integers(-1, 2, size=(part, N, 1)), -1)
        bad = rng.random((part, N)) < 0.05
        acts[bad] = rng.integers(-2, 3, size=(int(bad.sum()), ora.nd))
        want = ora.rollout(part, acts)
        got = env.rollout(part, actions=acts)
        for k in ("obs", "final_obs", "terminated", "truncated"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                                   rtol=1e-12, atol=1e-12)
    st = env.episode_stats()
    for k in ("episodes", "transitions", "noisy_transitions", "terminated"):
        assert st[k][0] == ora.stats[k], k
    np.testing.assert_allclose(st["reward"][0], ora.stats["reward"], rtol=1e-9)
    if ora.has_pnoise:  # the action was replaced about p of the (valid) time
        assert abs(ora.stats["noisy_transitions"] / ora.stats["transitions"]
                   - ora.p * 0.95) < 0.03


@pytest.mark.gpu
def test_ziggurat_normal_matches_oracle_and_differs_from_the_default():
    """normal_precision='ziggurat': numpy's Generator.normal algorithm on
    Philox words for the grid reward noise (the default here is fp64
    Box-Muller)."""
    cfg = gu.case_config("grid_sparse_noise")
    N, T = 1000, 24
    ora = VectorGridOracle(scalar_oracle(gu.case_config("grid_sparse_noise")), N,
                           autoreset=True, horizon=10, seed=5, normal="ziggurat")
    env = make_env(N, autoreset=True, horizon=10, philox_seed=5,
                   normal_precision="ziggurat", **cfg)
    ora.reset()
    acts = np.zeros((T, N, 2), dtype=np.int64)
    acts[..., 1] = np.random.default_rng(1).integers(-1, 2, size=(T, N))
    want, got = ora.rollout(T, acts), env.rollout(T, actions=acts)
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=1e-12, atol=1e-12)
    dflt = make_env(N, autoreset=True, horizon=10, philox_seed=5, **cfg)
    other = dflt.rollout(T, actions=acts)
    assert np.array_equal(other["obs"].cpu().numpy(), want["obs"])   # same transitions
    r = other["reward"].cpu().numpy()
    assert not np.allclose(r, want["reward"])
    from scipy import stats
    # sparse reward: r = (1[at target] + N(0, 1)) * 2 + 0.5 (+ 2 * 2 once terminal)
    z_zg = (want["reward"] - 0.5) / 2.0
    z_bm = (r - 0.5) / 2.0
    # the noise-free parts agree, so the difference of the two is a difference
    # of two independent N(0, 1): variance 2
    diff = (z_zg - z_bm).reshape(-1)
    assert stats.kstest(diff / np.sqrt(2.0), "norm").pvalue > 1e-3


@pytest.mark.gpu
def test_fast_normal_matches_oracle_within_1e5():
    """normal_precision='fast' (SFU Box-Muller): cells exact, rewards within
    1e-5 absolute of the oracle's fp32 restatement."""
    cfg = gu.case_config("grid_sparse_noise")
    N, T = 1000, 24
    ora = VectorGridOracle(scalar_oracle(gu.case_config("grid_sparse_noise")), N,
                           autoreset=True, horizon=10, seed=5, fast_normal=True)
    env = make_env(N, autoreset=True, horizon=10, philox_seed=5,
                   normal_precision="fast", **cfg)
    ora.reset()
    acts = np.zeros((T, N, 2), dtype=np.int64)
    acts[..., 1] = np.random.default_rng(1).integers(-1, 2, size=(T, N))
    want, got = ora.rollout(T, acts), env.rollout(T, actions=acts)
    for k in ("obs", "final_obs", "terminated", "truncated"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"],
                               rtol=0, atol=1e-5 * 2.0)


@pytest.mark.gpu
def test_graphed_step_and_rollout_equal_repeated_steps():
    import torch
    cfg = gu.case_config("grid_sparse_noise")
    a = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    b = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    c = make_env(512, autoreset=True, horizon=7, philox_seed=3, **cfg)
    step = b.make_graphed_step()
    gen = torch.Generator("cuda").manual_seed(0)
    acts = torch.zeros((12, 512, 2), dtype=torch.int64, device="cuda")
    acts[..., 0] = torch.randint(-1, 2, (12, 512), device="cuda", generator=gen)
    whole = c.rollout(12, actions=acts)
    for t in range(12):
        o1, r1, d1, t1, _ = a.step(acts[t])
        o2, r2, d2, t2, _ = step(acts[t])
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        assert torch.equal(o1, whole["obs"][t]) and torch.equal(r1, whole["reward"][t])


def test_grid_config_rejections():
    from mdp_playground_b200.config import parse_config
    base = gu.case_config("grid_dense_term")
    with pytest.raises(NotImplementedError):
        parse_config(dict(base, delay=1))
    with pytest.raises(NotImplementedError):
        parse_config(dict(base, sequence_length=2))
    cfg = dict(base)
    del cfg["make_denser"]
    with pytest.raises(ValueError):
        parse_config(cfg)


@pytest.mark.gpu
def test_augmented_state_round_trip():
    """get_augmented_state() mirrors the reference's [previous cell or NaN,
    current cell] window; set_augmented_state() on a fresh env continues the
    same trajectory."""
    import torch
    cfg = gu.case_config("grid_dense_term")
    ref = scalar_oracle(gu.case_config("grid_dense_term"))
    a = make_env(1, noise="numpy", **cfg)
    aug = a.get_augmented_state()["augmented_state"][0].cpu().numpy()
    assert np.isnan(aug[0]).all() and list(aug[1]) == list(ref.augmented_state[-1])
    for act in ([0, 1], [-1, 0], [1, 1], [0, -1]):
        ref.step(list(act))
        a.step(np.array([act]))
        aug = a.get_augmented_state()["augmented_state"][0].cpu().numpy()
        assert [list(x) for x in aug] == [list(x) for x in ref.augmented_state]
    b = make_env(1, noise="numpy", **gu.case_config("grid_dense_term"))
    b.set_augmented_state(a.get_augmented_state())
    for act in ([0, -1], [1, 0], [1, 0]):
        oa, ra, da, _, _ = a.step(np.array([act]))
        ob, rb, db, _, _ = b.step(np.array([act]))
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    b.set_augmented_state(torch.tensor([[2, 3]]))
    assert b.get_augmented_state()["curr_state"].tolist() == [[2, 3]]
    assert torch.isnan(b.get_augmented_state()["augmented_state"][0, 0]).all()


@pytest.mark.gpu
def test_action_substitution_distribution_chi2():
    """transition_noise = 1: the action is always replaced by a GridActionSpace
    sample different from it.  From an interior cell the displacement
    histogram must match the reference's rejection sampler: for a no-op the 4
    unit moves are equally likely; for a move, the other 3 moves have weight 1
    and the no-op weight 2 (two of the six samples are no-ops) out of 5.
    Chi-squared, threshold p > 1e-4 (3 / 3 dof: 21.1)."""
    import torch
    cfg = dict(gu.case_config("grid_sparse_noise"), transition_noise=1.0,
               grid_shape=(9, 9), target_point=[0, 0])
    N = 200_000
    for act, weights in (([0, 0], {(1, 0): 1, (-1, 0): 1, (0, 1): 1, (0, -1): 1}),
                         ([1, 0], {(0, 0): 2, (-1, 0): 1, (0, 1): 1, (0, -1): 1})):
        env = make_env(N, philox_seed=11, **cfg)
        start = torch.full((N, 2), 4, dtype=torch.int64)
        env.reset(options={"init_state": start})
        obs, *_ = env.step(torch.tensor([act]).repeat(N, 1))
        d = (obs.cpu().numpy() - 4)
        total = sum(weights.values())
        chi2 = 0.0
        seen = 0
        for move, w in weights.items():
            n = int(((d[:, 0] == move[0]) & (d[:, 1] == move[1])).sum())
            seen += n
            chi2 += (n - N * w / total) ** 2 / (N * w / total)
        assert seen == N, "a displacement outside the substitute set occurred"
        assert chi2 < 21.1, (act, chi2)
        st = env.episode_stats()
        assert st["noisy_transitions"][0] == N


@pytest.mark.gpu
def test_full_size_properties_1m_envs():
    """BASELINE-sized batch (1 M envs x 64 steps, no noise, no auto-reset):
    size-independent properties checked on the device -- every move follows
    the wall-bounce rule, the dense reward telescopes to the Manhattan progress,
    `terminated` is sticky from the first visit of the target."""
    import torch
    cfg = dict(gu.case_config("grid_dense_term"), reward_scale=1.0,
               term_state_reward=0.0)
    N, T = 1 << 20, 64
    env = make_env(N, philox_seed=1, track_history=False, **cfg)
    start = env.get_augmented_state()["curr_state"].clone()
    gen = torch.Generator("cuda").manual_seed(0)
    acts = torch.zeros((T, N, 2), dtype=torch.int64, device="cuda")
    dim = torch.randint(0, 2, (T, N), device="cuda", generator=gen)
    val = torch.randint(-1, 2, (T, N), device="cuda", generator=gen)
    acts.scatter_(2, dim[..., None], val[..., None])
    out = env.rollout(T, actions=acts)
    obs = out["obs"]
    prev = torch.cat([start[None], obs[:-1]], dim=0)
    want = (prev + acts).clamp_(min=0)
    want = torch.minimum(want, torch.tensor([7, 7], device="cuda"))
    assert torch.equal(obs, want)
    tgt = torch.tensor([5, 5], device="cuda")
    dist = (obs - tgt).abs().sum(-1)
    d0 = (start - tgt).abs().sum(-1)
    assert torch.equal(out["reward"].sum(0), (d0 - dist[-1]).to(torch.float64))
    at = (dist == 0)
    sticky = torch.cummax(at.to(torch.uint8), dim=0).values.bool()
    assert torch.equal(out["terminated"], sticky)
    assert not out["truncated"].any() and 0.3 < sticky[-1].float().mean() < 0.99
