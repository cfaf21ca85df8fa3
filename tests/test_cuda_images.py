"""GPU parity tests of the image-observation renderers: every pixel equal to
the reference's (Pillow) rendering -- against the reference's golden images,
the scalar oracle driven by the same numpy streams, and an exhaustive sweep
of the transform parameters."""
import warnings

import numpy as np
import pytest
import torch

from oracle.scalar_env import ScalarRLToyEnv
from tests import golden_util as gu
from tests.golden.cases import CASES
from tests.test_image_tables import oracle_image

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


@pytest.mark.parametrize("name", ["c4_img_shift", "c4_img_all", "img_none_64x48"])
def test_discrete_images_replay_reference_golden(name):
    g = gu.load(name)
    K, T = g["done"].shape
    env = make_env(K, noise="replay", **gu.case_config(name))
    obs, _ = env.reset(options={"reset_u": g["init_reset_u"],
                                "image_params": g["init_image_params"]})
    assert obs.shape == g["init_image"].shape and obs.dtype == torch.uint8
    assert np.array_equal(obs.cpu().numpy(), g["init_image"])
    for t in range(T):
        obs, r, term, trunc, info = env.step(
            g["actions"][:, t], replay=dict(
                transition_u=np.nan_to_num(g["transition_u"][:, t]),
                reward_noise=np.nan_to_num(g["reward_noise"][:, t]),
                image_params=g["image_params"][:, t]))
        assert np.array_equal(info["state"].cpu().numpy(), g["state"][:, t])
        assert np.array_equal(obs.cpu().numpy(), g["obs_image"][:, t]), t
        assert np.array_equal(r.cpu().numpy(), g["reward"][:, t])
        m = g["reset_after"][:, t]
        if m.any():
            obs, _ = env.reset(options={
                "mask": m, "reset_u": np.nan_to_num(g["reset_u"][:, t]),
                "image_params": np.where(m[:, None], g["reset_image_params"][:, t],
                                         g["image_params"][:, t])})
            assert np.array_equal(obs.cpu().numpy()[m], g["reset_image"][m, t]), t


@pytest.mark.parametrize("name", ["c4_img_shift", "c4_img_all"])
def test_discrete_images_same_seed_drop_in(name):
    """noise='numpy': the image stream (seed_dict['image_representations'])
    is consumed like the reference does => identical pixels, ctor included."""
    cfg = gu.case_config(name)
    ref = scalar_oracle(gu.case_config(name))
    env = make_env(1, noise="numpy", **cfg)
    assert np.array_equal(env.curr_obs[0].cpu().numpy(), ref.curr_obs)
    rng = np.random.default_rng(3)
    for t in range(60):
        a = int(rng.integers(8))
        o1, r1, d1, _, _ = ref.step(a)
        o2, r2, d2, _, _ = env.step([a])
        assert np.array_equal(o2[0].cpu().numpy(), o1), t
        assert float(r2[0]) == float(r1) and bool(d2[0]) == d1
        if d1 or t % 15 == 14:
            o1, _ = ref.reset()
            o2, _ = env.reset()
            assert np.array_equal(o2[0].cpu().numpy(), o1)


def test_discrete_renderer_exhaustive_parameter_sweep():
    """All 8 states x all 20 radii x sampled shifts / rotations / flips,
    8000 images, every pixel against Pillow."""
    cfg = gu.case_config("c4_img_all")
    env = make_env(8000, noise="replay", **cfg)
    tb = env.image_tables
    rng = np.random.default_rng(0)
    states = np.zeros(8000, dtype=np.int64)
    params = np.zeros((8000, 5), dtype=np.int32)
    i = 0
    for s in range(8):
        for R in range(tb.r_min, tb.r_min + tb.n_radii):
            m = 50 - R
            for _ in range(50):
                states[i] = s
                params[i] = (R, 50 + rng.integers(-m + 1, m),
                             50 + rng.integers(-m + 1, m),
                             rng.integers(-1, 360), rng.integers(3))
                i += 1
    assert i == 8000
    imgs = env.render_observation(torch.as_tensor(states, device="cuda"),
                                  image_params=params).cpu().numpy()
    E = type("E", (), dict(image_width=100, image_height=100))
    for k in range(8000):
        R, sw, sh, rot, flip = (int(v) for v in params[k])
        want = oracle_image(E, int(states[k]), dict(
            R=R, shift_w=sw, shift_h=sh, rotation=None if rot < 0 else rot,
            flip=flip))
        assert np.array_equal(imgs[k, :, :, 0], want), (k, params[k])


def test_discrete_images_philox_params_and_pixels():
    """Device-drawn transform parameters: legal ranges / quantisation, all
    radii and both flips occur, and the pixels match Pillow for the
    parameters the kernel reports."""
    cfg = dict(gu.case_config("c4_img_all"), image_sh_quant=4, image_ro_quant=15)
    N = 4096
    env = make_env(N, **cfg)
    out = env.rollout(3, want_final_obs=False)
    imgs = env.render_observation(out["obs"], step_index=env._step_index - 2)
    prm = env.last_image_params.cpu().numpy().reshape(3, N, 5)
    R, sw, sh, rot, flip = (prm[..., k] for k in range(5))
    assert R.min() == 10 and R.max() == 29 and len(np.unique(R)) == 20
    assert np.all((sw - 50) % 4 == 0) and np.all((sh - 50) % 4 == 0)
    # floor-quantisation can move a negative shift down by up to q-1
    assert np.all(sw - 50 < 50 - R) and np.all(sw - 50 >= -(50 - R) + 1 - 3)
    assert np.all(sh - 50 < 50 - R) and np.all(sh - 50 >= -(50 - R) + 1 - 3)
    assert np.all(rot % 15 == 0) and rot.min() == 0 and rot.max() == 345
    frac = [(flip == k).mean() for k in range(3)]
    assert abs(frac[0] - 0.5) < 0.03 and abs(frac[1] - 0.25) < 0.03
    # log-uniform radius: P(R < 17) = ln(17/10)/ln(3) = 0.483
    assert abs((R < 17).mean() - np.log(1.7) / np.log(3)) < 0.03
    # different steps / envs draw different parameters
    assert (prm[0] != prm[1]).any(axis=1).mean() > 0.99
    E = type("E", (), dict(image_width=100, image_height=100))
    st = out["obs"].cpu().numpy()
    im = imgs.cpu().numpy()
    for t in range(3):
        for n in range(0, N, 37):
            p = prm[t, n]
            want = oracle_image(E, int(st[t, n]), dict(
                R=int(p[0]), shift_w=int(p[1]), shift_h=int(p[2]),
                rotation=int(p[3]), flip=int(p[4])))
            assert np.array_equal(im[t, n, :, :, 0], want)
    # the per-step image API draws the same parameters as the block render
    env2 = make_env(N, **dict(cfg))
    o, _, _, _, info = env2.step(env2.action_space.sample(size=N))
    assert o.shape == (N, 100, 100, 1)


def test_continuous_images_replay_reference_golden():
    g = gu.load("cont_img")
    K, T = g["done"].shape
    env = make_env(K, noise="replay", **gu.case_config("cont_img"))
    obs, _ = env.reset(options={"init_state": g["init_state"]})
    assert np.array_equal(obs.cpu().numpy(), g["init_image"])
    for t in range(T):
        obs, r, term, trunc, info = env.step(torch.as_tensor(g["actions"][:, t]),
                                             replay={})
        assert np.array_equal(info["state"].cpu().numpy(), g["state"][:, t])
        assert np.array_equal(obs.cpu().numpy(), g["obs_image"][:, t]), t
        m = g["reset_after"][:, t]
        if m.any():
            obs, _ = env.reset(options={"mask": m, "init_state": g["reset_state"][:, t]})
            assert np.array_equal(obs.cpu().numpy()[m], g["reset_image"][m, t])


def test_continuous_renderer_random_states_vs_pillow():
    cfg = gu.case_config("cont_img")
    ref = scalar_oracle(gu.case_config("cont_img"))
    N = 3000
    env = make_env(N, **cfg)
    rng = np.random.default_rng(1)
    st = rng.uniform(-5, 5, size=(N, 4)).astype(np.float32)
    st[:50] = np.round(st[:50])          # pixel-boundary positions
    st[50:60] = [5.0, -5.0, 5.0, -5.0]   # corners: clipped stamps
    imgs = env.render_observation(torch.as_tensor(st, device="cuda")).cpu().numpy()
    assert imgs.shape == (N, 200, 100, 3)
    for k in range(0, N, 3):
        assert np.array_equal(imgs[k], ref._image_continuous(st[k])), k
    for k in range(60):
        assert np.array_equal(imgs[k], ref._image_continuous(st[k])), k


@pytest.mark.parametrize("W,H", [(50, 50), (60, 50), (50, 60), (64, 52), (200, 120)])
def test_discrete_renderer_other_image_sizes(W, H):
    """Sizes off the 100 x 100 default: 50 x 50 and 60 x 50 take the byte-wise
    path (W*H or H not a multiple of 16 / 4), the others the vector path with
    different strides; shift + rotate + flip, every pixel against Pillow."""
    cfg = dict(gu.case_config("c4_img_all"), image_width=W, image_height=H,
               image_transforms="shift,rotate,flip", image_sh_quant=1)
    cfg.pop("image_scale_range")
    n = 1200
    env = make_env(n, noise="replay", **cfg)
    rng = np.random.default_rng(W * 1000 + H)
    states = rng.integers(0, 8, size=n)
    R = 20
    mw, mh = W // 2 - R, H // 2 - R
    params = np.stack([np.full(n, R), W // 2 + rng.integers(-mw + 1, mw, size=n),
                       H // 2 + rng.integers(-mh + 1, mh, size=n),
                       rng.integers(-1, 360, size=n), rng.integers(3, size=n)],
                      axis=1).astype(np.int32)
    imgs = env.render_observation(torch.as_tensor(states, device="cuda"),
                                  image_params=params).cpu().numpy()
    assert imgs.shape == (n, W, H, 1)
    E = type("E", (), dict(image_width=W, image_height=H))
    for k in range(n):
        Rk, sw, sh, rot, flip = (int(v) for v in params[k])
        want = oracle_image(E, int(states[k]), dict(
            R=Rk, shift_w=sw, shift_h=sh, rotation=None if rot < 0 else rot,
            flip=flip))
        assert np.array_equal(imgs[k, :, :, 0], want), (k, params[k])


@pytest.mark.parametrize("transforms", ["shift", "shift,scale,rotate,flip"])
def test_buffered_image_step_equals_unbuffered(transforms):
    """step() of an image env through the pre-marshalled buffers (step kernel +
    renderer launched as its programmatic dependent) == the allocate-per-call
    path, pixel for pixel, incl. the irrelevant sub-image."""
    import torch
    for irr in (False, True):
        cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
                   state_space_size=[8, 8] if irr else 8,
                   action_space_size=[8, 8] if irr else 8,
                   irrelevant_features=irr, reward_density=0.25,
                   terminal_state_density=0.25, image_representations=True,
                   image_transforms=transforms, image_sh_quant=2, image_ro_quant=1,
                   image_scale_range=(0.5, 1.5))
        N = 257
        fast = make_env(N, autoreset=True, horizon=5, **cfg)
        slow = make_env(N, autoreset=True, horizon=5, step_buffers=0, **cfg)
        shape = (6, N, 2) if irr else (6, N)
        a = torch.randint(0, 8, shape, dtype=torch.int32, device="cuda",
                          generator=torch.Generator("cuda").manual_seed(4))
        for t in range(6):
            f, s = fast.step(a[t]), slow.step(a[t])
            assert f[0].shape == s[0].shape and f[0].dtype == torch.uint8
            assert torch.equal(f[0], s[0]), (irr, t)
            assert torch.equal(f[1], s[1]) and torch.equal(f[2], s[2])
            assert torch.equal(fast.last_image_params, slow.last_image_params)
