"""The committed fixtures ARE what the reference produces: regenerate every
golden case with tests/golden/make_golden.py (the unmodified reference under
the gymnasium stand-in) into a scratch directory and compare array by array
with the committed .npz files.  Needs the reference sources (/root/reference in
the build container, or the vendored oracle/_ref copy)."""
import warnings

import numpy as np
import pytest

from tests import golden_util as gu
from tests.golden.cases import CASES

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("name", sorted(CASES))
def test_fixture_regenerates_identically(name, tmp_path):
    from tests.golden import make_golden
    spec = CASES[name]
    grid = spec["config"]["state_space_type"] == "grid"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        (make_golden.run_grid_case if grid else make_golden.run_case)(
            name, spec, out_dir=str(tmp_path))
    fresh = dict(np.load(tmp_path / (name + ".npz"), allow_pickle=False))
    committed = gu.load(name)
    assert sorted(fresh) == sorted(committed)
    for k in committed:
        assert fresh[k].dtype == committed[k].dtype, k
        assert np.array_equal(fresh[k], committed[k], equal_nan=fresh[k].dtype.kind == "f"), k


def test_wrapper_fixtures_regenerate_identically(tmp_path):
    """Same for the GymEnvWrapper tail fixtures (tests/golden/wrap_*.npz)."""
    from tests.golden.make_wrapper_golden import WRAPPER_CASES, run_wrapper_case
    for name, spec in WRAPPER_CASES.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            run_wrapper_case(name, spec, out_dir=str(tmp_path))
        fresh = dict(np.load(tmp_path / (name + ".npz"), allow_pickle=False))
        committed = gu.load(name)
        assert sorted(fresh) == sorted(committed)
        for k in committed:
            assert np.array_equal(fresh[k], committed[k],
                                  equal_nan=fresh[k].dtype.kind == "f"), (name, k)
