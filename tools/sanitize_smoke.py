"""Small run of every kernel (AOT and NVRTC variants) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys, warnings
import numpy as np, torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
from tests import golden_util as gu
warnings.simplefilter("ignore")
for jit in (True, False):
    for name in ("c2_every1", "big50", "custom_8x5", "notmax_diam2"):
        env = VectorRLToyEnv(333, autoreset=True, horizon=7, **gu.case_config(name))
        env.set_jit(jit)
        env.rollout(21)                         # generic io, odd T
        a = torch.randint(0, env.tables.n_actions, (16, 333), dtype=torch.int32, device="cuda")
        env.rollout(16, actions=a, want_final_obs=False)   # fast io
        env.step(a[0]); env.reset(options={"mask": np.arange(333) % 2 == 0})
    het = VectorRLToyEnv(777, autoreset=True, horizon=5, config_groups=[gu.case_config(n) for n in ("c1_seq1", "c2_every1", "big50")])
    het.set_jit(jit); het.rollout(19)
    for name in ("c3_order2", "cont_noise_delay", "cont_term_boxes", "cont_order3"):
        cfg = gu.case_config(name)
        env = VectorRLToyEnv(301, autoreset=True, horizon=9, **cfg)
        env.set_jit(jit)
        D = cfg["state_space_dim"]
        env.rollout(13, torch.rand((13, 301, D), device="cuda") * 2 - 1)
        env.reset()
img = VectorRLToyEnv(65, autoreset=True, horizon=5, **gu.case_config("c4_img_all"))
for _ in range(3): img.step(torch.randint(0, 8, (65,), dtype=torch.int32, device="cuda"))
img2 = VectorRLToyEnv(33, **gu.case_config("img_none_64x48")); img2.step(torch.zeros(33, dtype=torch.int32, device="cuda"))
ci = VectorRLToyEnv(17, **gu.case_config("cont_img")); ci.step(torch.zeros((17, 4), device="cuda"))
torch.cuda.synchronize(); print("sanitize smoke done")
