#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched RLToyEnv step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): discrete toy MDP, 8 states x 8 actions,
sequence_length 3, delay 2, transition_noise 0.1, reward_noise 0.25,
65 536 envs per GPU, horizon 100 with auto-reset, native Philox noise.
One bench "step" = one fused rollout launch of --inner (default 1000)
env-steps over all envs of the rank: actions [T,N] int32 are read from HBM,
obs i64 / reward f64 / terminated u8 / truncated u8 [T,N] are written
(22 B per env-step, 1.44 GB per launch: larger than the 126 MB L2, so no
flush is needed between launches).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port of the
reference loop (oracle/scalar_env.py; the Python reference itself cannot
travel to the GPU box) on all host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ENVS_PER_GPU = 65536
ALGO_BYTES_ROLLOUT = 22   # action in + outputs out
ALGO_BYTES_STEP = 58      # SURVEY.md 8d table, config C2 single step


def workload_config():
    return dict(seed=0, state_space_type="discrete",
                action_space_type="discrete", state_space_size=8,
                action_space_size=8, sequence_length=3, delay=2,
                transition_noise=0.1, reward_noise=0.25, reward_density=0.25,
                terminal_state_density=0.25, generate_random_mdp=True,
                reward_every_n_steps=True)


WORKLOAD_NAME = ("discrete toy (dqn_p_r_noises/dqn_seq_del shape): 8 states, "
                 "8 actions, sequence_length=3, delay=2, transition_noise=0.1, "
                 "reward_noise=0.25, 65536 envs per GPU, horizon 100 auto-reset")


# --------------------------------------------------------------------------
# CPU port of the reference loop
# --------------------------------------------------------------------------
def _cpu_loop(n_steps, seed=0, horizon=100):
    """for t: env.step(a_t); reset on done or every `horizon` steps."""
    import numpy as np
    from oracle.scalar_env import ScalarRLToyEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = ScalarRLToyEnv(**dict(workload_config(), seed=seed))
    actions = np.random.default_rng(0xC0FFEE + seed).integers(
        0, 8, size=n_steps).tolist()
    t0 = time.perf_counter()
    since = 0
    for a in actions:
        _, _, done, _, _ = env.step(a)
        since += 1
        if done or since == horizon:
            env.reset()
            since = 0
    return time.perf_counter() - t0


def _cpu_worker(args):
    n_steps, seed = args
    return _cpu_loop(n_steps, seed)


def cpu_port_throughput(n_steps, procs):
    """Aggregate steps/s of `procs` independent scalar envs (1 = in-process)."""
    if procs == 1:
        dt = _cpu_loop(n_steps)
        return n_steps / dt
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(n_steps, s) for s in range(procs)])
        wall = time.perf_counter() - t0
    return procs * n_steps / wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    per_proc = args.ref_sample
    _cpu_loop(2000)  # import + warm caches
    for _ in range(args.warmup):
        cpu_port_throughput(max(per_proc // 10, 1000), cores)
    t0 = time.perf_counter()
    vals = [cpu_port_throughput(per_proc, cores) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    sample = (f"{cores} processes x {per_proc} env-steps of the scalar CPU "
              f"port per bench step, reset on done or every 100 steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "envs": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML,
    every 2 ms; nvidia-smi polling is too coarse for a 20-50 ms region)."""
    HW_SLOWDOWN, SW_POWER_CAP = 0x8, 0x4
    HW_THERMAL, SW_THERMAL = 0x40, 0x20

    def __init__(self, index):
        self.index, self.sm, self.reasons = index, [], 0
        self.stop_flag, self.thread, self.max_mhz = False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(
                self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = sorted(self.sm)
        names = [(self.HW_SLOWDOWN, "hw_slowdown"),
                 (self.HW_THERMAL, "hw_thermal_slowdown"),
                 (self.SW_THERMAL, "sw_thermal_slowdown"),
                 (self.SW_POWER_CAP, "sw_power_cap")]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in names if self.reasons & bit],
                "samples": len(sm)}


# --------------------------------------------------------------------------
def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n_envs, n_steps, jit):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
    kernel, from the committed `ncu --set full` capture (profiles/), when the
    launch shape matches the captured one; else null."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        k = t["mdpp_jit_rollout" if jit else "discrete_rollout_kernel"]
        if k["envs"] == n_envs and k["steps_per_launch"] == n_steps:
            return k["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def _time_launches(torch, fn, n, barrier, max_over_ranks):
    for _ in range(3):
        fn()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / n


def other_configs(torch, Env, dev, rank, world, barrier, max_over_ranks, peak):
    """The other BASELINE.json shapes, device-resident, one line each:
    env-steps/s (all ranks) and fraction of the HBM roofline with the
    algorithmic bytes of SURVEY.md 8d."""
    import numpy as np
    res = {}
    base = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
                state_space_size=8, action_space_size=8, reward_density=0.25,
                terminal_state_density=0.25)

    def line(name, n_envs, steps_per_launch, ms, bytes_per_step, note):
        sps = world * n_envs * steps_per_launch / (ms * 1e-3)
        res[name] = {"value": sps, "unit": UNIT, "envs_per_gpu": n_envs,
                     "env_steps_per_launch": steps_per_launch,
                     "ms_per_launch": ms,
                     "algorithmic_bytes_per_env_step": bytes_per_step,
                     "roofline_frac": sps / world * bytes_per_step / 1e9 / peak,
                     "what": note}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # C1: seq 1, delay 0, no noise
        N, T = 65536, 1000
        env = Env(N, device=dev, autoreset=True, horizon=100,
                  env_id_offset=rank * N, sequence_length=1, delay=0, **base)
        acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device=dev)
        out = env.rollout(T, actions=acts, want_final_obs=False)
        ms = _time_launches(torch, lambda: env.rollout(T, actions=acts, out=out),
                            10, barrier, max_over_ranks)
        line("C1_discrete_seq1_rollout", N, T, ms, 22,
             "fused rollout, noise off")
        del env, acts, out
        # N2: discrete irrelevant_features, dqn_irr_dims.py
        # shape; rows (relevant, irrelevant): 8 B actions in, 16 B obs out
        env = Env(N, device=dev, autoreset=True, horizon=100,
                  env_id_offset=rank * N, sequence_length=1, delay=0,
                  irrelevant_features=True, transition_noise=0.1,
                  **dict(base, state_space_size=[8, 8], action_space_size=[8, 8]))
        acts = torch.randint(0, 8, (T, N, 2), dtype=torch.int32, device=dev)
        out = env.rollout(T, actions=acts, want_final_obs=False)
        ms = _time_launches(torch, lambda: env.rollout(T, actions=acts, out=out),
                            10, barrier, max_over_ranks)
        line("N2_discrete_irrelevant_features_rollout", N, T, ms, 34,
             "fused rollout, two sub-MDPs per env, transition noise 0.1")
        del env, acts, out
        # N1: grid env, tests/test_mdp_playground.py:1057 shape;
        # int64 rows: 16 B action in, 16 B cell + 8 B reward + 2 B flags out
        Ng, Tg = 1 << 20, 100
        env = Env(Ng, device=dev, autoreset=True, horizon=100,
                  env_id_offset=rank * Ng, seed=0, state_space_type="grid",
                  grid_shape=(8, 8), delay=0, sequence_length=1,
                  reward_function="move_to_a_point", target_point=[5, 5],
                  make_denser=True, transition_noise=0.1)
        acts = torch.zeros((Tg, Ng, 2), dtype=torch.int64, device=dev)
        acts[..., 0] = torch.randint(-1, 2, (Tg, Ng), device=dev)
        out = env.rollout(Tg, actions=acts, want_final_obs=False)
        ms = _time_launches(torch, lambda: env.rollout(Tg, actions=acts, out=out),
                            5, barrier, max_over_ranks)
        line("N1_grid_rollout", Ng, Tg, ms, 42,
             "fused rollout, 8x8 grid, dense reward, transition noise 0.1")
        del env, acts, out
        # C3: continuous move_to_a_point, 1M envs
        N, T = 1 << 20, 100
        c3 = dict(seed=0, state_space_type="continuous", state_space_dim=6,
                  relevant_indices=[0, 1], irrelevant_features=True,
                  transition_dynamics_order=2, inertia=1.0, time_unit=0.5,
                  target_radius=0.05, target_point=[0.0, 0.0],
                  state_space_max=10.0, action_space_max=1.0)
        env = Env(N, device=dev, autoreset=True, horizon=100,
                  env_id_offset=rank * N, **c3)
        acts = torch.rand((T, N, 6), device=dev) * 2 - 1
        out = env.rollout(T, actions=acts, want_final_obs=False)
        ms = _time_launches(torch, lambda: env.rollout(T, actions=acts, out=out),
                            5, barrier, max_over_ranks)
        line("C3_continuous_rollout", N, T, ms, 54, "fused rollout, fp32")
        a1, o1 = acts[:1], {k: v[:1] for k, v in out.items()}
        ms = _time_launches(torch, lambda: env.rollout(1, actions=a1, out=o1),
                            50, barrier, max_over_ranks)
        line("C3_continuous_single_step", N, 1, ms, 256,
             "gym-style step(): one launch per step")
        gstep = env.make_graphed_step()
        gstep.actions.copy_(a1[0])  # actions resident in the graph's input buffer
        ms = _time_launches(torch, lambda: gstep(gstep.actions), 50, barrier,
                            max_over_ranks)
        line("C3_continuous_single_step_cuda_graph", N, 1, ms, 256,
             "step() replayed from a CUDA graph")
        del gstep
        del env, acts, out
        # C4: 100x100 image observations, 16384 envs
        N = 16384
        for tr, extra in (("shift", dict(image_sh_quant=4)),
                          ("shift,scale,rotate", dict(image_sh_quant=1,
                                                      image_ro_quant=1,
                                                      image_scale_range=(0.5, 1.5)))):
            env = Env(N, device=dev, autoreset=True, horizon=100,
                      env_id_offset=rank * N, sequence_length=1, delay=0,
                      image_representations=True, image_transforms=tr,
                      image_width=100, image_height=100, **extra, **base)
            a = torch.randint(0, 8, (N,), dtype=torch.int32, device=dev)
            ms = _time_launches(torch, lambda: env.step(a), 30, barrier,
                                max_over_ranks)
            line(f"C4_image_step[{tr}]", N, 1, ms, 10034,
                 "step() + render, two launches per step")
            gstep = env.make_graphed_step()
            gstep.actions.copy_(a)  # (resident in the graph's input buffer)
            ms = _time_launches(torch, lambda: gstep(gstep.actions), 30, barrier,
                                max_over_ranks)
            line(f"C4_image_step_cuda_graph[{tr}]", N, 1, ms, 10034,
                 "step() + render replayed from a CUDA graph")
            # fused form (SURVEY.md 8d ii): T steps in one rollout launch, all
            # T x N observations rendered by one launch
            Tr = 8
            acts = torch.randint(0, 8, (Tr, N), dtype=torch.int32, device=dev)
            out = env.rollout(Tr, actions=acts, want_final_obs=False)

            def rollout_and_render():
                env.rollout(Tr, actions=acts, out=out)
                return env.render_observation(out["obs"],
                                              step_index=env._step_index - Tr)
            ms = _time_launches(torch, rollout_and_render, 10, barrier,
                                max_over_ranks)
            line(f"C4_image_rollout[{tr}]", N, Tr, ms, 10014,
                 "rollout(8) + one render launch for the 8 x N observations")
            del env, gstep, acts, out
        # C5: 1000-cell heterogeneous grid, 1M envs per GPU
        cfgs = [dict(base, delay=d, sequence_length=L, transition_noise=pn,
                     reward_noise=rn, make_denser=md, reward_every_n_steps=True)
                for d in (0, 1, 2, 4, 8) for L in (1, 2, 3, 4)
                for pn in (0, 0.01, 0.02, 0.1, 0.25) for rn in (0, 1, 5, 10, 25)
                for md in (False, True)]
        N, T = 1 << 20, 100
        env = Env(N, device=dev, autoreset=True, horizon=100, config_groups=cfgs,
                  shard=(rank, world), normal_precision="fast")
        acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device=dev)
        out = env.rollout(T, actions=acts, want_final_obs=False)
        ms = _time_launches(torch, lambda: env.rollout(T, actions=acts, out=out),
                            5, barrier, max_over_ranks)
        line("C5_heterogeneous_1000_groups_rollout", N, T, ms, 22,
             "fused rollout, one multi-group launch (scalars shared by all groups specialised)")
        s5 = env.episode_stats(reduce=True)
        res["C5_heterogeneous_1000_groups_rollout"]["stats_allreduce"] = {
            "groups": len(cfgs),
            "episode_len_mean_min_max": [float(np.min(s5["episode_len_mean"])),
                                         float(np.max(s5["episode_len_mean"]))]}
    return res


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mdp_playground_b200 import VectorRLToyEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner (and any NCCL_DEBUG output) to the
        # process's stdout from C: keep file descriptor 1 for the ONE JSON
        # line of the bench contract and send everything else to stderr
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N, T = args.envs, args.inner

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = VectorRLToyEnv(N, device=dev, autoreset=True, horizon=100,
                             env_id_offset=rank * N,
                             normal_precision=args.normal, **workload_config())
    env.set_jit(not args.no_jit)
    # synthetic actions: Philox, seed 0xC0FFEE + rank
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xC0FFEE + rank)
    actions = torch.randint(0, 8, (T, N), dtype=torch.int32, device=dev,
                            generator=gen)
    out = {
        "obs": torch.empty((T, N), dtype=torch.int64, device=dev),
        "reward": torch.empty((T, N), dtype=torch.float64, device=dev),
        "terminated": torch.empty((T, N), dtype=torch.bool, device=dev),
        "truncated": torch.empty((T, N), dtype=torch.bool, device=dev),
    }

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: `value` and the roofline --------------------
    for _ in range(args.warmup):
        env.rollout(T, actions=actions, out=out)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        env.rollout(T, actions=actions, out=out)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * N * T * args.steps / (ms_total * 1e-3)
    per_gpu_steps_per_s = N * T / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = per_gpu_steps_per_s * ALGO_BYTES_ROLLOUT / 1e9

    # ---- gym-style single step() calls (launch-bound; reported) -
    n_single = 200
    a1 = actions[0:1]
    o1 = {k: v[0:1] for k, v in out.items()}
    for _ in range(10):
        env.rollout(1, actions=a1, out=o1)
    barrier()
    e0.record()
    for _ in range(n_single):
        env.rollout(1, actions=a1, out=o1)
    e1.record()
    barrier()
    single_ms = max_over_ranks(e0.elapsed_time(e1)) / n_single
    single_sps = world * N / (single_ms * 1e-3)
    # the same call replayed from a CUDA graph (env.make_graphed_step())
    gstep = env.make_graphed_step()
    gstep.actions.copy_(actions[0])  # actions resident in the graph's input buffer
    graph_ms = _time_launches(torch, lambda: gstep(gstep.actions), n_single, barrier,
                              max_over_ranks)
    graph_sps = world * N / (graph_ms * 1e-3)

    # ---- end-to-end leg: host buffers through the public API --------------
    h_act = torch.empty((T, N), dtype=torch.int32).pin_memory()
    h_act.copy_(actions.cpu())
    h_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory()
             for k, v in out.items()}

    def e2e_step():  # public host-buffer API: pipelined H2D / kernel / D2H
        env.rollout_host(T, h_act, h_out, chunk_steps=50)

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * N * T * e2e_steps / (e2e_ms * 1e-3)
    h2d = T * N * 4
    d2h = T * N * (8 + 8 + 1 + 1)

    # ---- end-of-run episode statistics: ONE all-reduce (NCCL) -------------
    summ = env.episode_stats(reduce=True)
    stats = {"episodes": float(summ["episodes"][0]),
             "transitions": float(summ["transitions"][0]),
             "episode_reward_mean": float(summ["episode_reward_mean"][0]),
             "episode_len_mean": float(summ["episode_len_mean"][0]),
             "noisy_transition_frac": float(summ["noisy_transitions"][0]
                                            / max(summ["transitions"][0], 1))}
    others = None if args.no_other_configs else other_configs(
        torch, VectorRLToyEnv, dev, rank, world, barrier, max_over_ranks, peak)

    if rank == 0:
        # ---- CPU baseline: scalar port, one core, bounded sample ----------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _cpu_loop(2000)
            n_cpu = args.cpu_sample
            v = cpu_port_throughput(n_cpu, 1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{n_cpu} env-steps of oracle/scalar_env.py "
                             "(CPU restatement of the reference loop), one "
                             "process, reset on done or every 100 steps"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "envs_per_gpu": N,
                       "env_steps_per_launch": T,
                       "noise": "philox4x32-10; reward normals: "
                                + ("fp32 SFU Box-Muller" if args.normal == "fast"
                                   else "fp64 Box-Muller"),
                       "kernel_build": "nvrtc-specialised" if env.jit_last_used
                       else "ahead-of-time",
                       "autoreset": "same-step", "l2_policy":
                       f"inputs+outputs {T * N * 22 / 1e6:.0f} MB per launch "
                       "> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(N, T, env.jit_last_used),
                         "peak_source": peak_src,
                         "kernel": "mdpp_jit_rollout" if env.jit_last_used
                         else "discrete_rollout_kernel<PHILOX,smem>",
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_ROLLOUT,
                         "env_steps_per_launch": N * T},
            "single_step_api": {"value": single_sps, "unit": UNIT,
                                "us_per_call": single_ms * 1e3,
                                "roofline_frac": single_sps / world
                                * ALGO_BYTES_STEP / 1e9 / peak,
                                "cuda_graph": {
                                    "value": graph_sps, "unit": UNIT,
                                    "us_per_call": graph_ms * 1e3,
                                    "roofline_frac": graph_sps / world
                                    * ALGO_BYTES_STEP / 1e9 / peak}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "episode_stats_allreduced": stats,
            "other_configs": others,
        }
        sys.stdout.flush()
        if world > 1:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU,
                    help="envs per GPU")
    ap.add_argument("--inner", type=int, default=1000,
                    help="env-steps fused into one launch (one bench step)")
    ap.add_argument("--cpu-sample", type=int, default=300000)
    ap.add_argument("--ref-sample", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the extra BASELINE.json shapes (C1/C3/C4/C5)")
    ap.add_argument("--normal", default="fast", choices=["fast", "fp64"],
                    help="reward-noise normals: SFU fp32 Box-Muller or fp64")
    ap.add_argument("--no-jit", action="store_true",
                    help="use the ahead-of-time kernels instead of the "
                         "NVRTC-specialised one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
