"""Time env.step() of an image env (step kernel + renderer per call).
    python tools/time_image_step.py [n_envs]"""
import sys
import warnings

import torch

sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
            state_space_size=8, action_space_size=8, reward_density=0.25,
            terminal_state_density=0.25, sequence_length=1, delay=0,
            image_representations=True, image_width=100, image_height=100)
for tr, extra in (("shift", dict(image_sh_quant=4)),
                  ("shift,scale,rotate", dict(image_sh_quant=1, image_ro_quant=1,
                                              image_scale_range=(0.5, 1.5)))):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, autoreset=True, horizon=100, image_transforms=tr,
                             **extra, **base)
    a = torch.randint(0, 8, (N,), dtype=torch.int32, device="cuda")
    for _ in range(5):
        env.step(a)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            env.step(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(f"image step [{tr}] N={N}: {best*1e3:.1f} us/step = "
          f"{N*10034/best/1e6/6534.1:.1%} of peak", flush=True)
