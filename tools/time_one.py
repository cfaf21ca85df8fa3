"""Time ONE variant of the headline (BASELINE configs[1]) rollout launch.
    python tools/time_one.py [fp64|boxmuller|fast] [n_envs] [T] [reps]
Used under ncu (`-k regex:mdpp_jit_rollout -s 3 -c 1`) and for A/B runs with
MDPP_JIT_EXTRA / MDPP_JIT_CHUNK / MDPP_JIT_MINBLOCKS."""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv  # noqa: E402

normal = sys.argv[1] if len(sys.argv) > 1 else "fp64"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
           state_space_size=8, action_space_size=8, sequence_length=3, delay=2,
           transition_noise=0.1, reward_noise=0.25, reward_density=0.25,
           terminal_state_density=0.25, reward_every_n_steps=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = VectorRLToyEnv(N, autoreset=True, horizon=100, normal_precision=normal, **cfg)
acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda")
out = env.rollout(T, actions=acts, want_final_obs=False)
for _ in range(3):
    env.rollout(T, actions=acts, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    env.rollout(T, actions=acts, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
sps = N * T / ms * 1e3
print(f"{normal} N={N} T={T}: {ms:.4f} ms {sps:.3e} steps/s frac {sps*22/1e9/6534.1:.3f} "
      f"jit={env.jit_last_used} {env.jit_log[:200]}", flush=True)
