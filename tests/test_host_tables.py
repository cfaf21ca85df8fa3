"""CPU tests of the product's host side: config defaults, seeded table
generation (must equal the reference's, via the golden files), and the C-ABI
library's exported symbols.  No GPU compute."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

from mdp_playground_b200 import _lib
from mdp_playground_b200.config import parse_config
from mdp_playground_b200.tables import build_discrete_tables
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", gu.DISCRETE_CASES)
def test_host_tables_equal_reference_golden(name):
    g = gu.load(name)
    cfg = gu.case_config(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = parse_config(cfg)
        tb = build_discrete_tables(sp)
    import json
    assert sp.seed_dict == json.loads(str(g["seed_dict"]))
    assert np.array_equal(tb.transition, g["P"])
    assert np.array_equal(tb.terminal_states, g["terminal_states"])
    assert np.array_equal(tb.init_state_dist, g["init_state_dist"])
    assert sp.reward_every_n_steps == int(g["reward_every_n_steps"])
    if not cfg.get("use_custom_mdp"):
        gold = gu.golden_sequences(g)
        assert list(tb.rewardable_sequences.items()) == list(gold.items())
        L = sp.sequence_length
        full = [(k, v) for k, v in gold.items() if len(k) == L]
        assert np.array_equal(tb.sequences, np.array([k for k, _ in full]))
        assert np.array_equal(tb.sequence_rewards, [v for _, v in full])


def test_survey_known_answer_seed0():
    """SURVEY.md 8c 'Extra KAT' (reference under seed=0, 8x8)."""
    sp = parse_config(dict(seed=0, state_space_type="discrete",
                           action_space_size=8, state_space_size=8,
                           sequence_length=3, delay=2))
    tb = build_discrete_tables(sp)
    assert sp.seed_dict["relevant_state_space"] == 5874934615388537134
    assert sp.seed_dict["image_representations"] == 5595227450766711102
    assert tb.transition.tolist()[:3] == [[0, 2, 4, 7, 1, 6, 5, 3],
                                          [6, 1, 5, 2, 7, 3, 4, 0],
                                          [6, 5, 3, 2, 4, 1, 0, 7]]
    assert tb.transition[6].tolist() == [6] * 8
    assert sorted(tb.terminal_states.tolist()) == [6, 7]
    assert list(tb.rewardable_sequences)[:5] == [
        (3, 4, 1), (1, 4, 3), (2, 3, 0), (3, 5, 4), (4, 5, 1)]
    assert len(tb.rewardable_sequences) == 30
    assert sp.reward_every_n_steps == 3


def test_defaults_and_rejections():
    sp = parse_config({})
    assert sp.kind == "discrete" and sp.state_space_size == 8
    assert sp.delay == 0 and sp.sequence_length == 1 and not sp.make_denser
    sp = parse_config(dict(state_space_type="Continuous", state_space_dim=3))
    assert sp.kind == "continuous" and sp.make_denser
    assert sp.relevant_indices == [0, 1, 2] and sp.reward_every_n_steps == 1
    with pytest.raises(AssertionError):
        parse_config(dict(state_space_type="discrete", action_space_size=[4, 4]))
    with pytest.raises(TypeError):
        parse_config(dict(state_space_type="discrete", action_space_size=4,
                          seed="x"))
    with pytest.raises(NotImplementedError):
        parse_config(dict(state_space_type="discrete", action_space_size=4,
                          reward_noise=lambda s, a, r: 0.0))
    with pytest.raises(ValueError):
        parse_config(dict(state_space_type="banana"))
    # unknown keys are ignored like the reference does
    parse_config(dict(state_space_type="discrete", action_space_size=4,
                      completely_connected=True, dummy_seed=3, dummy_eval=1))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mdpp_b200.h")).read()
    declared = set(re.findall(r"\b(mdpp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert os.path.exists(_lib.LIB_PATH), \
        "libmdpp_b200.so not built (python -m mdp_playground_b200.build)"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mdpp_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    # sizes computed from the header's field lists (LP64)
    assert ctypes.sizeof(_lib.DiscreteGroup) == 10 * 4 + 5 * 8 + 7 * 8 + 3 * 8 + 2 * 4 + 3 * 8
    assert ctypes.sizeof(_lib.DiscreteState) == 8 + 5 * 8 + 2 * 4 + 3 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.DiscreteIO) == 9 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.StepOpts) == 6 * 4 + 4 * 8


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mdp_playground_b200 import VectorRLToyEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VectorRLToyEnv(4, state_space_type="discrete", action_space_size=4)


def test_nvrtc_specialisation_compiles_without_a_gpu():
    """The embedded kernel sources must be NVRTC-clean (sm_100a)."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(1 << 16)
    rc = lib.mdpp_jit_selftest(buf, len(buf))
    assert rc == 0, buf.value.decode()
