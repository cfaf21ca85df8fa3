"""Where the multi-group (C5) launch loses against a single specialised group:
1 M envs x 100 steps with group sets that differ in ONE dimension at a time.
Run under gpurun."""
import sys, warnings
import torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8,
            action_space_size=8, reward_density=0.25, terminal_state_density=0.25,
            reward_every_n_steps=True)
one = dict(base, delay=2, sequence_length=3, transition_noise=0.1, reward_noise=1)
N, T = 1 << 20, 100


def run(name, cfgs, sizes=None):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if len(cfgs) == 1:
            env = VectorRLToyEnv(N, autoreset=True, horizon=100, normal_precision="fast", **cfgs[0])
        else:
            env = VectorRLToyEnv(N, autoreset=True, horizon=100, config_groups=cfgs,
                                 group_sizes=sizes, normal_precision="fast")
    acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda")
    out = env.rollout(T, actions=acts, want_final_obs=False)
    for _ in range(3): env.rollout(T, actions=acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): env.rollout(T, actions=acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10; sps = N * T / ms * 1e3
    print(f"{name:34s} groups={len(cfgs):5d} {ms:.4f} ms {sps:.3e} steps/s frac {sps*22/1e9/6534:.3f} "
          f"jit={env.jit_last_used}", flush=True)


run("single group", [one])
run("1000 identical groups", [one] * 1000)
run("1024 identical groups (aligned)", [one] * 1024)
run("1000 identical, sizes % 64 == 0", [one] * 1000, [1088] * 384 + [1024] * 616)
run("delay in 0,1,2,4,8", [dict(one, delay=d) for d in (0, 1, 2, 4, 8)] * 200)
run("delay in 0,1,2,4", [dict(one, delay=d) for d in (0, 1, 2, 4)] * 250)
run("L in 1..4", [dict(one, sequence_length=L) for L in (1, 2, 3, 4)] * 250)
run("L in 1..3", [dict(one, sequence_length=L) for L in (1, 2, 3)] * 333)
run("p-noise in 0..0.25", [dict(one, transition_noise=p) for p in (0, 0.01, 0.02, 0.1, 0.25)] * 200)
run("r-noise in 0..25", [dict(one, reward_noise=r) for r in (0, 1, 5, 10, 25)] * 200)
run("full C5 grid", [dict(base, delay=d, sequence_length=L, transition_noise=pn, reward_noise=rn,
                          make_denser=md)
                     for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4) for pn in (0, 0.01, 0.02, 0.1, 0.25)
                     for rn in (0, 1, 5, 10, 25) for md in (False, True)])
run("full C5 grid, sizes % 64 == 0",
    [dict(base, delay=d, sequence_length=L, transition_noise=pn, reward_noise=rn, make_denser=md)
     for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4) for pn in (0, 0.01, 0.02, 0.1, 0.25)
     for rn in (0, 1, 5, 10, 25) for md in (False, True)], [1088] * 384 + [1024] * 616)
