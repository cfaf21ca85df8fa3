import time, warnings, torch, numpy as np, sys
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8, action_space_size=8, sequence_length=3, delay=2, transition_noise=0.1, reward_noise=0.25, reward_density=0.25, terminal_state_density=0.25, reward_every_n_steps=True)
def timeit(N, T, cfg, actions_given=True, reps=5, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, **cfg, **kw)
    acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda") if actions_given else None
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(2): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(T, actions=acts, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    sps = N * T / ms * 1e3
    print(f"N={N} T={T} given={actions_given} noise={'transition_noise' in cfg}: {ms:.3f} ms  {sps:.3e} steps/s  {sps*22/1e9:.0f} GB/s(22B) ", flush=True)
timeit(65536, 1000, cfg)
timeit(65536, 1000, cfg, normal_precision="fast")
timeit(65536, 1000, cfg, actions_given=False)
c1 = {k: v for k, v in cfg.items() if k not in ("transition_noise", "reward_noise")}
timeit(65536, 1000, c1)
timeit(65536, 1000, dict(c1, sequence_length=1, delay=0))
timeit(1 << 20, 200, cfg)
timeit(1 << 20, 200, c1)
timeit(65536, 1, cfg, reps=20)
