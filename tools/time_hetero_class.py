"""What per-class launches of the C5 grid could buy: the same 1 M envs x 100
steps with every group sharing one delay (and reward noise on / off)."""
import sys, warnings
import torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8,
            action_space_size=8, reward_density=0.25, terminal_state_density=0.25)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
N, T = 1 << 20, 100
for delays, rns, Ls in (((0, 1, 2, 4, 8), (0, 1, 5, 10, 25), (1, 2, 3, 4)), ((2,), (1, 5, 10, 25), (1, 2, 3, 4)),
                        ((2,), (1, 5, 10, 25), (3,)), ((8,), (1, 5, 10, 25), (1, 2, 3, 4)), ((0,), (0,), (1, 2, 3, 4))):
    cfgs = [dict(base, delay=d, sequence_length=L, transition_noise=pn, reward_noise=rn, make_denser=md,
                 reward_every_n_steps=True)
            for d in delays for L in Ls for pn in (0, 0.01, 0.02, 0.1, 0.25)
            for rn in rns for md in (False, True)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, config_groups=cfgs, normal_precision=prec)
    acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda")
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(3): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): env.rollout(T, actions=acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10; sps = N * T / ms * 1e3
    print(f"{prec} delays={delays} rn={rns} L={Ls} groups={len(cfgs)}: {ms:.4f} ms frac {sps*22/1e9/6534.1:.3f}", flush=True)
    del env, acts, out
