"""C5 (1000 heterogeneous config groups, 1 M envs) rollout timing; run under gpurun."""
import sys, warnings
import torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8,
            action_space_size=8, reward_density=0.25, terminal_state_density=0.25)
cfgs = [dict(base, delay=d, sequence_length=L, transition_noise=pn, reward_noise=rn, make_denser=md,
             reward_every_n_steps=True)
        for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4) for pn in (0, 0.01, 0.02, 0.1, 0.25)
        for rn in (0, 1, 5, 10, 25) for md in (False, True)]
N, T = 1 << 20, (int(sys.argv[2]) if len(sys.argv) > 2 else 100)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = VectorRLToyEnv(N, autoreset=True, horizon=100, config_groups=cfgs, normal_precision=(sys.argv[1] if len(sys.argv) > 1 else "fast"))
acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda")
out = env.rollout(T, actions=acts, want_final_obs=False)
for _ in range(3): env.rollout(T, actions=acts, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): env.rollout(T, actions=acts, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10; sps = N * T / ms * 1e3
print(f"C5 {sys.argv[1:]} T={T}: {ms:.4f} ms {sps:.3e} steps/s frac {sps*22/1e9/6534:.3f} jit={env.jit_last_used}", flush=True)
