"""Device time of the image renderer alone: K render launches captured in a
CUDA graph (no Python / launch overhead between them), timed with CUDA events.
    python tools/time_render.py [N]
"""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv  # noqa: E402

base = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
            state_space_size=8, action_space_size=8, sequence_length=1, delay=0,
            reward_density=0.25, terminal_state_density=0.25,
            image_representations=True, image_width=100, image_height=100)
PEAK = 6534.1


def timeit(N, cfg, K=10, reps=5):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, **cfg)
    st = torch.randint(0, 8, (N,), device="cuda")
    for _ in range(3):
        env.render_observation(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.render_observation(st)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for k in range(K):
            img = env.render_observation(st, step_index=k)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    gbs = N * 10000 / best / 1e6
    print(f"render N={N} {cfg.get('image_transforms'):>26}: {best*1e3:7.1f} us/launch "
          f"{gbs:6.0f} GB/s = {gbs/PEAK:.1%} of measured peak", flush=True)


N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
timeit(N, dict(base, image_transforms="none"))
timeit(N, dict(base, image_transforms="shift", image_sh_quant=4))
timeit(N, dict(base, image_transforms="shift,scale,rotate",
               image_scale_range=(0.5, 1.5), image_ro_quant=1, image_sh_quant=4))
timeit(N, dict(base, image_transforms="shift,scale,rotate,flip",
               image_scale_range=(0.5, 1.5)))
