"""CPU validation of the renderer tables and gather algorithm against Pillow
(through the scalar oracle, which renders exactly like the reference)."""
import warnings

import numpy as np
import pytest

from mdp_playground_b200 import image_tables as it
from oracle import render_emulation as emu
from oracle.scalar_env import ReplayDraws, ScalarRLToyEnv
from tests import golden_util as gu


def oracle_image(env, state, prm):
    """Render with the oracle (PIL) for explicit transform parameters."""
    import PIL.Image as Image
    import PIL.ImageDraw as ImageDraw
    sides = state + 3
    img = Image.new("L", (env.image_width, env.image_height))
    pts = [(int(prm["shift_w"] + prm["R"] * np.cos((2 * np.pi / sides) * i)),
            int(prm["shift_h"] + prm["R"] * np.sin((2 * np.pi / sides) * i)))
           for i in range(sides)]
    ImageDraw.Draw(img).polygon(pts, fill=255)
    if prm["rotation"] is not None:
        img = img.rotate(prm["rotation"])
    if prm["flip"] == 1:
        img = img.transpose(Image.FLIP_LEFT_RIGHT)
    elif prm["flip"] == 2:
        img = img.transpose(Image.FLIP_TOP_BOTTOM)
    return np.array(img).T


@pytest.fixture(scope="module")
def tables_all():
    return it.build_discrete_image_tables(8, 100, 100, "shift,scale,rotate,flip",
                                          2, 1, (0.5, 1.5))


def test_rotation_tables_match_pillow_all_angles():
    it.rotation_coefficients(100, 100, check=True)
    it.rotation_coefficients(64, 48, check=True)
    it.rotation_coefficients(37, 91, check=True)


def test_disc_stamp_is_the_97_pixel_ellipse():
    spans = it.disc_stamp(5)
    assert spans[:, 1].tolist() == [5, 7, 9, 11, 11, 11, 11, 11, 9, 7, 5]
    assert spans[:, 1].sum() == 97
    assert np.array_equal(spans[:, 0] * 2 + spans[:, 1], np.ones(11))


def test_atlas_is_small(tables_all):
    tb = tables_all
    assert tb.r_min == 10 and tb.n_radii == 20
    assert tb.n_xvar <= 8 and tb.n_yvar <= 8
    assert tb.mask_bits.shape[0] < 4000


def test_emulation_equals_pillow_exhaustive_shifts(tables_all):
    """Every state x radius x a sweep of shifts, no rotation: all pixels."""
    tb = tables_all
    env = type("E", (), dict(image_width=100, image_height=100))
    rng = np.random.default_rng(0)
    n = 0
    for s in range(8):
        for R in range(tb.r_min, tb.r_min + tb.n_radii):
            m = 50 - R
            for _ in range(6):
                sw = 50 + int(rng.integers(-m + 1, m))
                sh = 50 + int(rng.integers(-m + 1, m))
                prm = dict(R=R, shift_w=sw, shift_h=sh, rotation=None, flip=0)
                want = oracle_image(env, s, prm)
                got = emu.render_discrete(tb, s, R, sw, sh, -1, 0)
                assert np.array_equal(got, want), (s, R, sw, sh)
                n += 1
    assert n == 8 * 20 * 6


def test_emulation_equals_pillow_random_full_transforms(tables_all):
    tb = tables_all
    env = type("E", (), dict(image_width=100, image_height=100))
    rng = np.random.default_rng(1)
    for _ in range(600):
        s = int(rng.integers(8))
        R = int(rng.integers(tb.r_min, tb.r_min + tb.n_radii))
        m = 50 - R
        sw = 50 + int(rng.integers(-m + 1, m))
        sh = 50 + int(rng.integers(-m + 1, m))
        rot = int(rng.integers(360))
        flip = int(rng.integers(3))
        prm = dict(R=R, shift_w=sw, shift_h=sh, rotation=rot, flip=flip)
        assert np.array_equal(emu.render_discrete(tb, s, R, sw, sh, rot, flip),
                              oracle_image(env, s, prm)), prm


def test_radius_thresholds_equal_reference_expression(tables_all):
    tb = tables_all
    rng = np.random.default_rng(2)
    us = np.concatenate([rng.random(20000), tb.r_thresholds,
                         np.nextafter(tb.r_thresholds, 0), [0.0]])
    lo, hi = np.log(0.5 * 20), np.log(1.5 * 20)
    for u in us:
        assert emu.radius_from_uniform(tb, u) == int(np.exp(lo + u * (hi - lo)))


@pytest.mark.parametrize("name", ["c4_img_shift", "c4_img_all", "img_none_64x48"])
def test_emulation_reproduces_reference_golden_images(name):
    """Reference images (golden) from the recorded transform parameters."""
    g = gu.load(name)
    cfg = gu.case_config(name)
    tb = it.build_discrete_image_tables(
        8, cfg.get("image_width", 100), cfg.get("image_height", 100),
        cfg.get("image_transforms", "none"), cfg.get("image_sh_quant"),
        cfg.get("image_ro_quant"), cfg.get("image_scale_range"))
    K, T = g["done"].shape
    for k in range(K):
        for t in range(T):
            R, sw, sh, rot, flip = (int(v) for v in g["image_params"][k, t])
            got = emu.render_discrete(tb, int(g["state"][k, t]), R, sw, sh, rot, flip)
            assert np.array_equal(got, g["obs_image"][k, t, :, :, 0]), (k, t)
