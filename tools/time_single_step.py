import sys, warnings, torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
cfg = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8, action_space_size=8, sequence_length=3, delay=2, transition_noise=0.1, reward_noise=0.25, reward_density=0.25, terminal_state_density=0.25, reward_every_n_steps=True)
N = 65536
def t(env, label):
    a = torch.randint(0, 8, (1, N), dtype=torch.int32, device="cuda")
    out = env.rollout(1, actions=a, want_final_obs=False)
    for _ in range(20): env.rollout(1, actions=a, out=out)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): env.rollout(1, actions=a, out=out)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(20): env.rollout(1, actions=a, out=out)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(label, f"{e0.elapsed_time(e1)/20*1e3:.2f} us per T=1 launch (20 in a graph)", flush=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = VectorRLToyEnv(N, autoreset=True, horizon=100, normal_precision="fast", **cfg)
    t(env, "with stats   ")
    env._state.stats = None
    t(env, "stats = NULL ")
