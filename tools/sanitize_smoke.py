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
# ---- round 2 ---------------------------------------------------------------
from mdp_playground_b200 import VectorGymEnvTail
for jit in (True, False):
    for prec in ("fp64", "boxmuller", "fast"):           # staged ziggurat windows, queue, slow pass
        env = VectorRLToyEnv(333, autoreset=True, horizon=40, normal_precision=prec,
                             **gu.case_config("c2_every1"))
        env.set_jit(jit)
        a = torch.randint(0, 8, (75, 333), dtype=torch.int32, device="cuda")
        env.rollout(75, actions=a, want_final_obs=False)      # 2 windows + direct tail
        env.rollout(3, actions=a[:3], want_final_obs=False)   # peel, short launch (PDL)
        env.rollout(70, actions=a[:70])                       # generic signature, staged
        for t in range(4): env.step(a[t])                     # buffered step(), PDL chain
    het = VectorRLToyEnv(40 * 37, autoreset=True, horizon=23,
                         config_groups=[dict(gu.case_config("c2_every1"), delay=d, sequence_length=L,
                                             reward_noise=rn)
                                        for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4) for rn in (0, 5)],
                         group_sizes=[37] * 40)
    het.set_jit(jit)
    het.rollout(75, actions=torch.randint(0, 8, (75, 40 * 37), dtype=torch.int32, device="cuda"))
    for dt in (np.uint8, np.int32):                           # dtype_o written in-kernel
        env = VectorRLToyEnv(129, autoreset=True, horizon=9, dtype_o=dt, **gu.case_config("c2_every1"))
        env.set_jit(jit); env.rollout(17); env.step(torch.zeros(129, dtype=torch.int32, device="cuda"))
for name in ("cont_line_seq10", "cont_line_seq3_delay", "cont_line_2d", "cont_inertia_list",
             "cont_inertia_f32", "cont_delay_default_target"):
    cfg = gu.case_config(name)
    for jit in (True, False):
        env = VectorRLToyEnv(301, autoreset=True, horizon=19, **cfg)
        env.set_jit(jit)
        D = cfg["state_space_dim"]
        env.rollout(33, torch.rand((33, 301, D), device="cuda") * 2 - 1)
        env.step(torch.zeros((301, D), device="cuda")); env.reset()
cells = [dict(gu.case_config("c3_order2"), time_unit=tu, delay=d, transition_dynamics_order=o,
              reward_noise=0.1)
         for tu in (0.1, 0.5) for d in (0, 3) for o in (1, 2, 3)]
het = VectorRLToyEnv(12 * 100, autoreset=True, horizon=11, config_groups=cells, group_sizes=[100] * 12)
het.rollout(25, torch.rand((25, 1200, 6), device="cuda") * 2 - 1); het.reset()
tail = VectorGymEnvTail(1000, n_actions=6, state_space_type="discrete", delay=9, transition_noise=0.2,
                        reward_noise=0.5, reward_scale=2.0)
for t in range(12):
    a = tail.actions(torch.randint(0, 6, (1000,), dtype=torch.int32, device="cuda"))
    tail.post(None, torch.rand(1000, device="cuda"), torch.rand(1000, device="cuda") < 0.1)
tc = VectorGymEnvTail(500, obs_dim=3, state_space_type="continuous", delay=2, transition_noise=0.1,
                      reward_noise=0.1)
for t in range(5): tc.post(torch.rand((500, 3), device="cuda"), torch.rand(500, device="cuda"),
                           torch.zeros(500, dtype=torch.bool, device="cuda"))
ti = VectorGymEnvTail(33, image_side=12, state_space_type="discrete", image_transforms="shift",
                      image_padding=5, image_sh_quant=2)
ti.shift_images(torch.randint(0, 255, (33, 12, 12, 3), dtype=torch.uint8, device="cuda"))
torch.cuda.synchronize(); print("sanitize smoke done")
