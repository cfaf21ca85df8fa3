"""GymEnvWrapper's post-processing tail: the CPU
restatement (oracle/wrapper_tail.py) and the batched CUDA tail
(mdp_playground_b200.VectorGymEnvTail) against golden vectors recorded from the
reference wrapper (tests/golden/wrap_*.npz): bit-exact on every non-terminal
step; terminal steps (where the reference raises, gym_env_wrapper.py:414) are
checked CUDA-vs-oracle only."""
import numpy as np
import pytest

from oracle.wrapper_tail import ScalarWrapperTail
from tests import golden_util as gu
from tests.golden.make_wrapper_golden import WRAPPER_CASES


def replay_wrapper_golden(name, make_tail):
    """Drive a batched tail implementation (K lanes) through a golden case;
    returns (rewards [K, T], raised [K, T]) for further checks."""
    g = gu.load(name)
    out = replay_wrapper_record(g, WRAPPER_CASES[name], make_tail)
    assert g["raised"].sum() > 10 and (~g["raised"]).sum() > 150
    return out


def replay_wrapper_record(g, spec, make_tail):
    """The same for any record made by make_wrapper_golden.run_wrapper_case
    (the fuzz tests record fresh ones)."""
    K, T = g["raised"].shape
    tail = make_tail(K, spec)
    got_r = np.zeros((K, T))
    for t in range(T):
        applied = tail.actions(g["action"][:, t], g["choice_u"][:, t])
        assert np.array_equal(np.asarray(applied), g["applied"][:, t]), t
        obs, r = tail.post(g["base_obs"][:, t], g["base_reward"][:, t],
                           g["base_done"][:, t], g["reward_noise"][:, t],
                           g["obs_noise"][:, t], g["shift"][:, t])
        ok = ~g["raised"][:, t]
        assert np.array_equal(np.asarray(r)[ok], g["out_reward"][ok, t]), t
        assert np.array_equal(np.asarray(obs)[ok], g["out_obs"][ok, t]), t
        got_r[:, t] = r
    return got_r, g["raised"]


class OracleLanes:
    def __init__(self, K, spec):
        self.tails = [ScalarWrapperTail(n_actions=spec.get("n_actions"),
                                        **spec["config"]) for _ in range(K)]

    def actions(self, a, u):
        return [t.action(x if np.ndim(x) else int(x), ui)
                for t, x, ui in zip(self.tails, a, u)]

    def post(self, obs, r, done, rn, on, sh):
        out_o, out_r = [], []
        for i, t in enumerate(self.tails):
            out_o.append(t.observation(obs[i], None if np.isnan(on[i]).all() else on[i],
                                       sh[i]))
            out_r.append(t.reward(float(r[i]), bool(done[i]),
                                  None if np.isnan(rn[i]) else rn[i]))
        return np.array(out_o), np.array(out_r)


@pytest.mark.parametrize("name", sorted(WRAPPER_CASES))
def test_oracle_replays_reference_wrapper_golden(name):
    replay_wrapper_golden(name, OracleLanes)


class CudaLanes:
    def __init__(self, K, spec):
        import torch
        from mdp_playground_b200 import VectorGymEnvTail
        self.torch = torch
        cfg = spec["config"]
        self.tail = VectorGymEnvTail(
            K, noise="replay", n_actions=spec.get("n_actions"),
            obs_dim=spec.get("dim"), image_side=spec.get("side"), **cfg)
        self.cont = cfg["state_space_type"] == "continuous"

    def actions(self, a, u):
        if self.cont:
            return a
        return self.tail.actions(a, replay_u=u).cpu().numpy()

    def post(self, obs, r, done, rn, on, sh):
        o, rr = self.tail.post(obs, r, done, replay_reward_noise=rn,
                               replay_obs_noise=on if self.cont else None,
                               replay_shift=sh)
        o = o.cpu().numpy() if self.torch.is_tensor(o) else np.asarray(o)
        return o, rr.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(WRAPPER_CASES))
def test_cuda_tail_replays_reference_wrapper_golden(name):
    """Bit-exact against the reference wrapper on every non-terminal step; on
    the terminal steps (reference: TypeError) equal to the CPU restatement."""
    got, raised = replay_wrapper_golden(name, CudaLanes)
    want, _ = replay_wrapper_golden(name, OracleLanes)
    assert np.array_equal(got, want)
    assert raised.any()


@pytest.mark.gpu
def test_cuda_tail_philox_statistics():
    """Native noise: the substituted actions are uniform over the others with
    the configured probability (chi-squared), the reward noise is N(0, sigma)
    (KS), shifts are uniform on the quantised grid, delays hold."""
    import torch
    from scipy import stats
    from mdp_playground_b200 import VectorGymEnvTail
    N, A = 200_000, 6
    tail = VectorGymEnvTail(N, n_actions=A, seed=5, state_space_type="discrete",
                            delay=3, transition_noise=0.3, reward_noise=2.0,
                            reward_scale=1.0)
    a = torch.full((N,), 2, dtype=torch.int32, device="cuda")
    out = tail.actions(a).cpu().numpy()
    counts = np.bincount(out, minlength=A)
    expected = np.full(A, 0.3 / (A - 1) * N)
    expected[2] = 0.7 * N
    assert stats.chisquare(counts, expected).pvalue > 1e-4
    zeros = torch.zeros(N, dtype=torch.float64, device="cuda")
    nd = torch.zeros(N, dtype=torch.bool, device="cuda")
    outs = []
    for t in range(5):
        _, r = tail.post(None, zeros + (t + 1), nd)
        outs.append(r.cpu().numpy())
    assert stats.kstest(outs[0] / 2.0, "norm").pvalue > 1e-4   # delayed: pure noise
    assert abs(outs[3].mean() - 1.0) < 0.02 and abs(outs[4].mean() - 2.0) < 0.02
    img = VectorGymEnvTail(4096, seed=1, image_side=8, state_space_type="discrete",
                           image_transforms="shift", image_padding=6, image_sh_quant=2)
    x = torch.full((4096, 8, 8, 3), 200, dtype=torch.uint8, device="cuda")
    y = img.shift_images(x)
    assert y.shape == (4096, 20, 20, 3)
    assert int((y > 0).sum()) == 4096 * 8 * 8 * 3       # every image pasted whole
    sh = img.last_shift.cpu().numpy()
    assert sh.min() == -5 and sh.max() == 5
    assert stats.chisquare(np.bincount(sh[:, 0] + 5, minlength=11)).pvalue > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("side,pad,quant", [
    (84, 20, 2),   # Atari frames: shared-memory kernel, word-wise tile load
    (10, 5, 1),    # rows of 30 bytes: byte-wise tile load
    (12, 7, 3),    # quantised shifts, even canvas
    (300, 4, 2),   # tile larger than shared memory: the byte-wise kernel
    (16, 1, 1),    # one pixel of padding: the only shift is 0
])
def test_cuda_image_shift_equals_the_oracle_for_every_kernel_path(side, pad, quant):
    """Random images and the kernel's own (Philox) shifts against the CPU
    restatement of get_transformed_image (gym_env_wrapper.py:523-618)."""
    import torch
    from mdp_playground_b200 import VectorGymEnvTail
    from oracle.wrapper_tail import ScalarWrapperTail
    N = 64 if side < 100 else 6
    cfg = dict(state_space_type="discrete", image_transforms="shift",
               image_padding=pad, image_sh_quant=quant)
    tail = VectorGymEnvTail(N, seed=3, n_actions=4, image_side=side, **cfg)
    ora = ScalarWrapperTail(n_actions=4, **cfg)
    imgs = torch.randint(0, 256, (N, side, side, 3), dtype=torch.uint8, device="cuda",
                         generator=torch.Generator("cuda").manual_seed(side))
    got = tail.shift_images(imgs).cpu().numpy()
    shifts = tail.last_shift.cpu().numpy()
    lo, hi = ora.shift_draw_bounds(side)
    assert shifts.min() >= lo and shifts.max() < hi
    x = imgs.cpu().numpy()
    for i in range(N):
        assert np.array_equal(got[i], ora.shift_image(x[i], shifts[i])), i


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cuda_tail_observation_noise_statistics(dtype):
    """Native observation noise of continuous envs (:367-376): every dimension
    N(0, sigma) (KS), dimensions and steps uncorrelated, odd obs_dim handled;
    without transition_noise the observations pass through."""
    import torch
    from scipy import stats
    from mdp_playground_b200 import VectorGymEnvTail
    N, D = 100_000, 5
    dt = getattr(torch, dtype)
    tail = VectorGymEnvTail(N, seed=11, obs_dim=D, obs_dtype=dt, state_space_type="continuous",
                            transition_noise=0.5, reward_noise=1.0)
    base = torch.arange(D, dtype=dt, device="cuda").repeat(N, 1)
    r = torch.zeros(N, dtype=torch.float64, device="cuda")
    d = torch.zeros(N, dtype=torch.bool, device="cuda")
    o1, r1 = tail.post(base, r, d)
    o2, _ = tail.post(base, r, d)
    assert o1.dtype == dt and o1.shape == (N, D)
    z1 = ((o1 - base) / 0.5).double().cpu().numpy()
    z2 = ((o2 - base) / 0.5).double().cpu().numpy()
    for k in range(D):
        assert stats.kstest(z1[:, k], "norm").pvalue > 1e-4, k
    cc = np.corrcoef(np.concatenate([z1, z2], axis=1).T)
    assert np.abs(cc - np.eye(2 * D)).max() < 0.02
    assert stats.kstest(r1.cpu().numpy(), "norm").pvalue > 1e-4
    quiet = VectorGymEnvTail(N, seed=11, obs_dim=D, obs_dtype=dt, state_space_type="continuous")
    o3, r3 = quiet.post(base, r + 2.0, d)
    assert torch.equal(o3, base) and torch.equal(r3, r + 2.0)
