import warnings, torch, numpy as np, sys
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
cfg = dict(seed=0, state_space_type="continuous", action_space_type="continuous", state_space_dim=6, action_space_dim=6, relevant_indices=[0, 1], irrelevant_features=True, transition_dynamics_order=2, inertia=1.0, time_unit=0.5, target_radius=0.05, target_point=[0.0, 0.0], state_space_max=10.0, action_space_max=1.0, reward_function="move_to_a_point")
def timeit(N, T, cfg, reps=5, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, **cfg, **kw)
    acts = (torch.rand((T, N, 6), device="cuda") * 2 - 1)
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(2): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(T, actions=acts, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts)); sps = N * T / ms * 1e3
    bytes_step = 54 if T > 1 else 256
    print(f"cont N={N} T={T} noise={'transition_noise' in cfg} jit={env.jit_last_used}: {ms:.3f} ms {sps:.3e} steps/s {sps*bytes_step/1e9:.0f} GB/s({bytes_step}B)", flush=True)
timeit(1 << 20, 100, cfg)
timeit(1 << 20, 1, cfg, reps=20)
timeit(1 << 20, 100, dict(cfg, transition_noise=0.05, reward_noise=0.1))
timeit(1 << 20, 100, dict(cfg, transition_noise=0.05, reward_noise=0.1), normal_precision="boxmuller")
timeit(1 << 20, 100, dict(cfg, transition_noise=0.05, reward_noise=0.1), normal_precision="fast")
timeit(1 << 22, 50, cfg)
