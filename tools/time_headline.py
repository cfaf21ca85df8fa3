import warnings, torch, numpy as np, sys
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8, action_space_size=8, sequence_length=3, delay=2, transition_noise=0.1, reward_noise=0.25, reward_density=0.25, terminal_state_density=0.25, reward_every_n_steps=True)
N, T = 65536, 1000
for name, c, kw in (("C2 fast", cfg, dict(normal_precision="fast")), ("C2 fp64 ziggurat", cfg, {}), ("C2 fp64 boxmuller", cfg, dict(normal_precision="boxmuller")), ("C1", dict({k: v for k, v in cfg.items() if k not in ("transition_noise", "reward_noise")}, sequence_length=1, delay=0), {})):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, **c, **kw)
    acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda")
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(3): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): env.rollout(T, actions=acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20; sps = N * T / ms * 1e3
    print(f"{name}: {ms:.4f} ms {sps:.3e} steps/s frac {sps*22/1e9/6534:.3f} jit={env.jit_last_used}", flush=True)
