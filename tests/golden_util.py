"""Helpers shared by the golden-vector tests (oracle and CUDA side)."""
import os

import numpy as np

from tests.golden.cases import CASES, LANE_SEED, materialise

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

IRR_CASES = [k for k, v in CASES.items()
             if v["config"]["state_space_type"] == "discrete"
             and v["config"].get("irrelevant_features")]
DISCRETE_CASES = [k for k, v in CASES.items()
                  if v["config"]["state_space_type"] == "discrete"
                  and k not in IRR_CASES]
GRID_CASES = [k for k, v in CASES.items()
              if v["config"]["state_space_type"] == "grid"]
CONTINUOUS_CASES = [k for k, v in CASES.items()
                    if v["config"]["state_space_type"] == "continuous"]


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"),
                        allow_pickle=False))


def case_config(name):
    return materialise(CASES[name]["config"])


def golden_sequences(g):
    """rewardable_sequences dict (insertion order) from a golden file."""
    out, pos = {}, 0
    for n, r in zip(g["seq_len"], g["seq_reward"]):
        out[tuple(int(x) for x in g["seq_flat"][pos:pos + n])] = float(r)
        pos += n
    return out


def lane_seed(k):
    return LANE_SEED + 17 * k
