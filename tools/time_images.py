import warnings, torch, numpy as np, sys
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8, action_space_size=8, sequence_length=1, delay=0, reward_density=0.25, terminal_state_density=0.25, image_representations=True, image_width=100, image_height=100)
def timeit(N, cfg, reps=20):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, **cfg)
    st = torch.randint(0, 8, (N,), device="cuda")
    for _ in range(3): env.render_observation(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): img = env.render_observation(st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"render N={N} {cfg.get('image_transforms')}: {ms*1e3:.1f} us/launch  {N/ms*1e3:.3e} img/s  {N*10000/ms/1e6:.0f} GB/s", flush=True)
    acts = torch.randint(0, 8, (N,), dtype=torch.int32, device="cuda")
    for _ in range(3): env.step(acts)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): env.step(acts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  step()+render: {ms*1e3:.1f} us  {N/ms*1e3:.3e} steps/s  {N*10034/ms/1e6:.0f} GB/s(10034B)", flush=True)
timeit(16384, dict(base, image_transforms="shift", image_sh_quant=4))
timeit(16384, dict(base, image_transforms="shift,scale,rotate", image_scale_range=(0.5, 1.5), image_ro_quant=1, image_sh_quant=1))
timeit(65536, dict(base, image_transforms="shift,scale,rotate,flip", image_scale_range=(0.5, 1.5)))
