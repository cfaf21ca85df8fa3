"""Time the grid rollout launch: one configuration vs. config groups.
    python tools/time_grid.py [n_envs] [T] [n_groups]"""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
G = int(sys.argv[3]) if len(sys.argv) > 3 else 128
base = dict(seed=0, state_space_type="grid", grid_shape=(8, 8), target_point=[5, 5],
            delay=0, sequence_length=1, reward_function="move_to_a_point",
            make_denser=True, transition_noise=0.25, reward_noise=1.0)


def time_env(env, label):
    gen = torch.Generator("cuda").manual_seed(0)
    acts = torch.zeros((T, N, 2), dtype=torch.int64, device="cuda")
    dim = torch.randint(0, 2, (T, N, 1), device="cuda", generator=gen)
    acts.scatter_(2, dim, torch.randint(-1, 2, (T, N, 1), device="cuda", generator=gen))
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(3):
        env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        env.rollout(T, actions=acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    sps = N * T / ms * 1e3
    print(f"{label} N={N} T={T}: {ms:.4f} ms {sps:.3e} steps/s frac {sps*42/1e9/6534.1:.3f}",
          flush=True)


with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    time_env(VectorRLToyEnv(N, autoreset=True, horizon=100, **base), "single")
    cells = [dict(base, transition_noise=0.05 * (g % 5), reward_noise=0.5 * (g % 3),
                  make_denser=bool(g & 1)) for g in range(G)]
    time_env(VectorRLToyEnv(N, autoreset=True, horizon=100, config_groups=cells),
             f"{G} groups")
