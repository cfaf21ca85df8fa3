"""20 step() calls of the headline config (for an ncu launch list)."""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv  # noqa: E402

normal = sys.argv[1] if len(sys.argv) > 1 else "fp64"
cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
           state_space_size=8, action_space_size=8, sequence_length=3, delay=2,
           transition_noise=0.1, reward_noise=0.25, reward_density=0.25,
           terminal_state_density=0.25, reward_every_n_steps=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = VectorRLToyEnv(65536, autoreset=True, horizon=100, normal_precision=normal, **cfg)
a = torch.randint(0, 8, (65536,), dtype=torch.int32, device="cuda")
for _ in range(20):
    env.step(a)
torch.cuda.synchronize()
