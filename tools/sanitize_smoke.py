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
    for name in ("irr_8x8_noise", "irr_6x10_diam2"):   # irrelevant sub-MDP, rows of 2
        cfg = gu.case_config(name)
        env = VectorRLToyEnv(333, autoreset=True, horizon=7, **cfg)
        env.set_jit(jit)
        env.rollout(21)
        a = torch.stack([torch.randint(0, n, (16, 333), dtype=torch.int32, device="cuda")
                         for n in cfg["action_space_size"]], dim=-1)
        env.rollout(16, actions=a, want_final_obs=False)
        env.step(a[0]); env.reset(options={"mask": np.arange(333) % 2 == 0})
for name in ("grid_sparse_noise", "grid_irr_5x9"):      # grid kernels, 2 and 4 dims
    cfg = gu.case_config(name)
    env = VectorRLToyEnv(301, autoreset=True, horizon=9, **cfg)
    nd = env._nd
    a = torch.randint(-1, 3, (23, 301, nd), device="cuda")
    env.rollout(23, actions=a); env.step(a[0]); env.reset(options={"mask": np.arange(301) % 3 == 0})
for name in ("irr_img_all", "grid_img", "grid_img_irr"):  # stacked sub-images, grid images
    cfg = gu.case_config(name)
    env = VectorRLToyEnv(37, autoreset=True, horizon=5, **cfg)
    if cfg["state_space_type"] == "grid":
        a = torch.zeros((37, env._nd), dtype=torch.int64, device="cuda"); a[:, 0] = 1
    else:
        a = torch.randint(0, 8, (37, 2), dtype=torch.int32, device="cuda")
    for _ in range(3): env.step(a)
    g = env.make_graphed_step()                         # programmatic dependent launch path
    for _ in range(3): g(a)
img = VectorRLToyEnv(65, autoreset=True, horizon=5, **gu.case_config("c4_img_all"))
for _ in range(3): img.step(torch.randint(0, 8, (65,), dtype=torch.int32, device="cuda"))
img2 = VectorRLToyEnv(33, **gu.case_config("img_none_64x48")); img2.step(torch.zeros(33, dtype=torch.int32, device="cuda"))
ci = VectorRLToyEnv(17, **gu.case_config("cont_img")); ci.step(torch.zeros((17, 4), device="cuda"))
torch.cuda.synchronize(); print("sanitize smoke done")
