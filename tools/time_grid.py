"""Grid-env rollout timing (1 M envs x 100 steps); run under gpurun."""
import sys, warnings
import torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
N, T = 1 << 20, 100
for extra in (dict(), dict(transition_noise=0.1), dict(transition_noise=0.1, reward_noise=0.5),
              dict(transition_noise=0.1, reward_noise=0.5, _fast=True)):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fast = extra.pop("_fast", False)
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, seed=0, state_space_type="grid",
                             normal_precision="fast" if fast else "fp64",
                             grid_shape=(8, 8), delay=0, sequence_length=1,
                             reward_function="move_to_a_point", target_point=[5, 5],
                             make_denser=True, **extra)
    acts = torch.zeros((T, N, 2), dtype=torch.int64, device="cuda")
    acts[..., 0] = torch.randint(-1, 2, (T, N), device="cuda")
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(3): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env.rollout(T, actions=acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5; sps = N * T / ms * 1e3
    print(f"grid {extra}: {ms:.3f} ms {sps:.3e} steps/s {sps*42/1e9:.0f} GB/s frac {sps*42/1e9/6534:.3f}", flush=True)
