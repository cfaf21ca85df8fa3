"""Randomised configurations: the UNMODIFIED reference (through
tests/golden/make_golden.py, i.e. the same recorder that wrote the committed
fixtures) against the scalar oracle, beyond the fixed golden cases.

Every case draws a configuration from a seeded generator (fuzz_configs.py),
records a few short lanes of the reference into a scratch directory and checks
the oracle on them with both legs of tests/test_oracle_golden.py: its own
PCG64 streams (states, rewards, tables bit for bit) and the recorded draws.
Needs the reference sources, so it runs in the build container only; the GPU
side of the same configurations is tests/test_cuda_fuzz.py (CUDA vs oracle).
"""
import warnings

import numpy as np
import pytest

from tests.fuzz_configs import (continuous_fuzz_config, discrete_fuzz_config,
                                grid_fuzz_config, image_fuzz_config,
                                DISCRETE_SEEDS, CONTINUOUS_SEEDS, GRID_SEEDS,
                                IMAGE_SEEDS)
from tests.golden.cases import materialise
from tests.test_oracle_golden import numpy_leg, replay_leg

pytestmark = pytest.mark.reference


def _record(name, cfg, tmp_path, lanes, steps, horizon, grid=False):
    from tests.golden import make_golden
    spec = dict(config=cfg, lanes=lanes, steps=steps, horizon=horizon)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        (make_golden.run_grid_case if grid else make_golden.run_case)(
            name, spec, out_dir=str(tmp_path))
    return dict(np.load(tmp_path / (name + ".npz"), allow_pickle=False))


@pytest.mark.parametrize("seed", DISCRETE_SEEDS)
def test_discrete_fuzz_reference_vs_oracle(seed, tmp_path):
    cfg = discrete_fuzz_config(seed)
    g = _record("fuzz_d%d" % seed, cfg, tmp_path, lanes=3, steps=40, horizon=9)
    numpy_leg(g, materialise(cfg), 9)
    replay_leg(g, lambda: materialise(cfg), 9)


@pytest.mark.parametrize("seed", CONTINUOUS_SEEDS)
def test_continuous_fuzz_reference_vs_oracle(seed, tmp_path):
    cfg = continuous_fuzz_config(seed)
    g = _record("fuzz_c%d" % seed, cfg, tmp_path, lanes=3, steps=40, horizon=11)
    numpy_leg(g, materialise(cfg), 11)
    replay_leg(g, lambda: materialise(cfg), 11)


@pytest.mark.parametrize("seed", GRID_SEEDS)
def test_grid_fuzz_reference_vs_oracle(seed, tmp_path):
    from tests.test_grid import grid_numpy_leg
    cfg = grid_fuzz_config(seed)
    g = _record("fuzz_g%d" % seed, cfg, tmp_path, lanes=3, steps=40, horizon=9,
                grid=True)
    grid_numpy_leg(g, materialise(cfg), 9)


@pytest.mark.parametrize("seed", IMAGE_SEEDS)
def test_image_fuzz_reference_vs_oracle(seed, tmp_path):
    """Image observations (Pillow on the reference side, the oracle's
    restatement of its rasteriser on ours): every pixel of every step."""
    cfg = image_fuzz_config(seed)
    g = _record("fuzz_i%d" % seed, cfg, tmp_path, lanes=2, steps=16, horizon=6)
    numpy_leg(g, materialise(cfg), 6)


@pytest.mark.parametrize("seed", __import__("tests.fuzz_configs", fromlist=["x"]).WRAPPER_SEEDS)
def test_wrapper_tail_fuzz_reference_vs_oracle(seed, tmp_path):
    """The reference GymEnvWrapper around the stand-in base envs, recorded with
    the recipe of the wrap_*.npz fixtures, against the oracle's tail on every
    non-terminal step (the reference raises on terminal ones, :414)."""
    from tests.fuzz_configs import wrapper_fuzz_spec
    from tests.golden.make_wrapper_golden import run_wrapper_case
    from tests.test_wrapper_tail import OracleLanes, replay_wrapper_record
    spec = wrapper_fuzz_spec(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        run_wrapper_case("fuzz_w%d" % seed, spec, out_dir=str(tmp_path), lanes=3, steps=40)
    g = dict(np.load(tmp_path / ("fuzz_w%d.npz" % seed), allow_pickle=False))
    replay_wrapper_record(g, spec, OracleLanes)
    assert (~g["raised"]).sum() > 60
