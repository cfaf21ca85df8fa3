"""Generate golden vectors by running the UNMODIFIED reference RLToyEnv.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case in `cases.py` it builds the reference env, then drives
K independent "lanes" (the env re-seeded per lane) with scripted random
actions and a reset policy (on `done`, or every H steps), recording per step
the action, emitted state, reward, done flag, image observation, and every
random draw the step path consumed (through oracle/draw_recorder.py), so the
trajectories can be replayed bit-exactly by the oracle and the CUDA kernels.
Output: tests/golden/<case>.npz (compressed; a few KB each).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import draw_recorder  # noqa: E402
from oracle.ref_loader import make_reference_env  # noqa: E402
from tests.golden.cases import CASES, LANE_SEED, materialise  # noqa: E402


def _take(log, tag, kind):
    """Pop the draws of one step from the recorder log."""
    vals = [e for e in log if e[0] == tag and e[1] == kind]
    return vals


def run_case(name, spec, out_dir=HERE):
    cfg = materialise(spec["config"])
    K, T, H = spec.get("lanes", 6), spec.get("steps", 48), spec.get("horizon", 12)
    env = make_reference_env(cfg)
    cont = cfg["state_space_type"] == "continuous"
    image = bool(cfg.get("image_representations", False))
    irr = (not cont) and bool(cfg.get("irrelevant_features", False))
    D = cfg.get("state_space_dim", 0)
    out = {}
    if not cont:
        out["P"] = np.array(env.transition_matrix, dtype=np.int64)
        if not cfg.get("use_custom_mdp"):
            keys = list(env.rewardable_sequences.keys())
            out["seq_len"] = np.array([len(k) for k in keys], dtype=np.int64)
            out["seq_flat"] = np.array([s for k in keys for s in k],
                                       dtype=np.int64)
            out["seq_reward"] = np.array(
                [env.rewardable_sequences[k] for k in keys], dtype=np.float64)
        out["terminal_states"] = np.array(env.config["terminal_states"],
                                          dtype=np.int64)
        out["init_state_dist"] = np.array(
            env.config["relevant_init_state_dist"], dtype=np.float64)
    if irr:
        out["P_irr"] = np.array(env.config["transition_function_irrelevant"],
                                dtype=np.int64)
    out["seed_dict"] = np.array(json.dumps(env.seed_dict))
    out["reward_every_n_steps"] = np.array(int(env.reward_every_n_steps))

    sshape = (K, T, D) if cont else ((K, T, 2) if irr else (K, T))
    sdtype = np.float32 if cont else np.int64
    rec = dict(
        actions=np.zeros(sshape, dtype=np.float32 if cont else np.int64),
        state=np.zeros(sshape, dtype=sdtype),
        reward=np.zeros((K, T), dtype=np.float64),
        reward_is_f32=np.zeros((K, T), dtype=bool),
        done=np.zeros((K, T), dtype=bool),
        reset_after=np.zeros((K, T), dtype=bool),
        reset_state=np.zeros(sshape, dtype=sdtype),
        init_state=np.zeros((K, D) if cont else ((K, 2) if irr else (K,)),
                            dtype=sdtype),
        transition_u=np.full((K, T), np.nan),
        reward_noise=np.full((K, T), np.nan),
        reset_u=np.full((K, T), np.nan),
        init_reset_u=np.full((K,), np.nan),
    )
    if cont:
        rec["state_noise"] = np.full((K, T, D), np.nan)
        # (a list / float64 `inertia` makes the top derivative float64, :1654)
        f64_top = np.asarray(cfg.get("inertia", 1.0)).dtype == np.float64 \
            and np.ndim(cfg.get("inertia", 1.0)) > 0
        rec["derivs"] = np.zeros(
            (K, T, env.dynamics_order + 1, D),
            dtype=np.float64 if f64_top else np.float32)
    if irr:  # draws of the irrelevant sub-space (S' stream / second E draw)
        rec["irr_transition_u"] = np.full((K, T), np.nan)
        rec["irr_reset_u"] = np.full((K, T), np.nan)
        rec["init_irr_reset_u"] = np.full((K,), np.nan)
    n_sub = 2 if irr else 1
    pshape = (n_sub, 5) if irr else (5,)
    if image:
        shp = env.curr_obs[0].shape if isinstance(env.curr_obs, tuple) \
            else env.curr_obs.shape
        rec["obs_image"] = np.zeros((K, T) + shp, dtype=np.uint8)
        rec["init_image"] = np.zeros((K,) + shp, dtype=np.uint8)
        rec["reset_image"] = np.zeros((K, T) + shp, dtype=np.uint8)
        if not cont:
            # R, shift_w, shift_h, rotation(-1 = none), flip(0/1 LR/2 TB)
            rec["image_params"] = np.full((K, T) + pshape, -1, dtype=np.int64)
            rec["init_image_params"] = np.full((K,) + pshape, -1, dtype=np.int64)
            rec["reset_image_params"] = np.full((K, T) + pshape, -1,
                                                dtype=np.int64)

    def image_params_from(log_slice):
        """Decode the I-stream draws of one generate_image call."""
        tr = cfg.get("image_transforms", "none")
        W, Ht = cfg.get("image_width", 100), cfg.get("image_height", 100)
        it = iter([e for e in log_slice if e[0] == "image"])
        res = [one_image_params(it, tr, W, Ht) for _ in range(n_sub)]
        assert next(it, None) is None
        return res if irr else res[0]

    def one_image_params(it, tr, W, Ht):
        R = 20
        sw, sh, rot, flip = int(W / 2), int(Ht / 2), -1, 0
        if "scale" in tr:
            lo, hi = cfg.get("image_scale_range", (0.5, 1.5))
            u = float(next(it)[3])
            R = int(np.exp(np.log(lo * 20) + u * (np.log(hi * 20)
                                                   - np.log(lo * 20))))
        if "shift" in tr:
            q = cfg.get("image_sh_quant", 1)
            sw += (int(next(it)[3]) // q) * q
            sh += (int(next(it)[3]) // q) * q
        if "rotate" in tr:
            q = cfg.get("image_ro_quant", 1)
            rot = (int(next(it)[3]) // q) * q
        if "flip" in tr:
            if int(next(it)[3]) == 0:
                flip = 1 if int(next(it)[3]) == 0 else 2
        return [R, sw, sh, rot, flip]

    for k in range(K):
        lane_seed = LANE_SEED + 17 * k
        # re-seed every stream the step path draws from
        if not cont:
            env.observation_spaces[0].seed(lane_seed + 1)
            if irr:
                env.observation_spaces[1].seed(lane_seed + 4)
        else:
            env.feature_space.seed(lane_seed + 2)
        if image:
            env.observation_space.seed(lane_seed + 3)
        log = draw_recorder.install(env)
        obs0, _ = env.reset(seed=lane_seed)  # E replaced: re-wrap it below
        if not cont:
            g0 = np.random.Generator(np.random.PCG64(
                np.random.SeedSequence(lane_seed)))
            rec["init_reset_u"][k] = g0.random()
            if irr:
                rec["init_irr_reset_u"][k] = g0.random()
            rec["init_state"][k] = env.curr_state
        else:
            rec["init_state"][k] = env.curr_state
        if image:
            rec["init_image"][k] = obs0
            if not cont:
                rec["init_image_params"][k] = image_params_from(log)
        env._np_random = draw_recorder.RecordingGenerator(
            env._np_random, log, "env")
        arng = np.random.default_rng(1000 + k)
        for t in range(T):
            del log[:]
            if cont:
                amax = cfg.get("action_space_max", 1.0)
                # ~5 % of the actions fall outside the action space
                a = arng.uniform(-1.05 * amax, 1.05 * amax, size=D).astype(
                    np.float32)
            elif irr:
                a = np.array([arng.integers(n) for n in env.action_space_size])
            else:
                a = int(arng.integers(env.action_space_size[0]))
            rec["actions"][k, t] = a
            obs, r, done, trunc, info = env.step(a)
            rec["state"][k, t] = env.curr_state
            rec["reward"][k, t] = float(r)
            rec["reward_is_f32"][k, t] = isinstance(r, np.float32)
            rec["done"][k, t] = done
            for e in log:
                if e[1] == "choice_u" and e[0] == "obs0":
                    rec["transition_u"][k, t] = e[2]
                elif e[1] == "choice_u" and e[0] == "obs1":
                    rec["irr_transition_u"][k, t] = e[2]
                elif e[1] == "normal" and np.ndim(e[3]) == 0:
                    rec["reward_noise"][k, t] = float(e[3])
                elif e[1] == "normal":
                    rec["state_noise"][k, t] = e[3]
            if cont:
                rec["derivs"][k, t] = np.array(env.state_derivatives)
            if image:
                rec["obs_image"][k, t] = obs
                if not cont:
                    rec["image_params"][k, t] = image_params_from(log)
            if done or t % H == H - 1:
                del log[:]
                rec["reset_after"][k, t] = True
                obs_r, _ = env.reset()
                rec["reset_state"][k, t] = env.curr_state
                us = [e[2] for e in log if e[1] == "choice_u" and e[0] == "env"]
                if us:  # (continuous resets draw from the feature space, not E)
                    rec["reset_u"][k, t] = us[0]
                    if irr:
                        rec["irr_reset_u"][k, t] = us[1]
                if image:
                    rec["reset_image"][k, t] = obs_r
                    if not cont:
                        rec["reset_image_params"][k, t] = image_params_from(log)
    out.update(rec)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return out


def run_grid_case(name, spec, out_dir=HERE):
    """Grid envs (rl_toy_env.py:1727-1778, :1947-1965, :2325-2345): actions are
    unit moves, the draws are the E-stream uniform that decides whether the
    action is replaced, the accepted replacement (GridActionSpace.sample, A
    stream), the reward normal and the reset cell (F stream)."""
    cfg = materialise(spec["config"])
    K, T, H = spec.get("lanes", 6), spec.get("steps", 48), spec.get("horizon", 12)
    env = make_reference_env(cfg)
    nd = len(env.grid_shape)
    image = bool(cfg.get("image_representations", False))
    out = {"seed_dict": np.array(json.dumps(env.seed_dict)),
           "reward_every_n_steps": np.array(int(env.reward_every_n_steps))}
    rec = dict(
        actions=np.zeros((K, T, nd), dtype=np.int64),
        state=np.zeros((K, T, nd), dtype=np.int64),
        reward=np.zeros((K, T), dtype=np.float64),
        done=np.zeros((K, T), dtype=bool),
        reset_after=np.zeros((K, T), dtype=bool),
        reset_state=np.zeros((K, T, nd), dtype=np.int64),
        init_state=np.zeros((K, nd), dtype=np.int64),
        grid_noise_u=np.full((K, T), np.nan),
        grid_noise_action=np.zeros((K, T, nd), dtype=np.int64),  # the action applied
        grid_noise_attempts=np.zeros((K, T), dtype=np.int64),
        reward_noise=np.full((K, T), np.nan))
    if image:
        shp = env.curr_obs[0].shape
        rec["obs_image"] = np.zeros((K, T) + shp, dtype=np.uint8)
        rec["init_image"] = np.zeros((K,) + shp, dtype=np.uint8)
        rec["reset_image"] = np.zeros((K, T) + shp, dtype=np.uint8)
    attempts_log = []
    for k in range(K):
        lane_seed = LANE_SEED + 17 * k
        env.feature_space.seed(lane_seed + 2)
        env.action_space.seed(lane_seed + 5)
        log = draw_recorder.install(env)
        obs0, _ = env.reset(seed=lane_seed)
        rec["init_state"][k] = env.curr_state
        if image:
            rec["init_image"][k] = obs0
        env._np_random = draw_recorder.RecordingGenerator(env._np_random, log, "env")
        arng = np.random.default_rng(1000 + k)
        for t in range(T):
            del log[:]
            a = [0] * nd
            kind = arng.integers(12)
            if kind < 5:
                a[arng.integers(nd)] = int(arng.integers(-1, 2))
            elif kind < 10:   # a greedy move, so that targets do get reached
                d = int(arng.integers(2))
                a[d] = int(np.sign(env.target_point[d] - env.curr_state[d]))
            elif kind == 10:  # invalid: two moves at once -> noop
                a = [int(x) for x in arng.integers(-1, 2, size=nd)]
            else:             # invalid: out of range -> noop
                a[0] = 2
            rec["actions"][k, t] = a
            obs, r, done, trunc, info = env.step(list(a))
            rec["state"][k, t] = env.curr_state
            rec["reward"][k, t] = float(r)
            rec["done"][k, t] = done
            applied = list(a)
            acts = [e for e in log if e[0] == "action"]
            for e in log:
                if e[0] == "env" and e[1] == "uniform":
                    rec["grid_noise_u"][k, t] = float(e[3])
                elif e[0] == "env" and e[1] == "normal":
                    rec["reward_noise"][k, t] = float(e[3])
            if acts:  # (ind, val) pairs; the last pair is the accepted one
                pairs = [(int(acts[i][3]), int(acts[i + 1][3]) - 1)
                         for i in range(0, len(acts), 2)]
                attempts_log.append((k, t, pairs))
                rec["grid_noise_attempts"][k, t] = len(pairs)
                applied = [0] * nd
                applied[pairs[-1][0]] = pairs[-1][1]
            rec["grid_noise_action"][k, t] = applied
            if image:
                rec["obs_image"][k, t] = obs
            # odd lanes keep stepping after `done` (reached_terminal is sticky)
            if (done and k % 2 == 0) or t % H == H - 1:
                rec["reset_after"][k, t] = True
                obs_r, _ = env.reset()
                rec["reset_state"][k, t] = env.curr_state
                if image:
                    rec["reset_image"][k, t] = obs_r
    # every attempt, flattened, for the scalar oracle's replay leg
    flat = [(k, t, i, v) for k, t, ps in attempts_log for i, v in ps]
    out["grid_attempts"] = np.array(flat, dtype=np.int64).reshape(-1, 4)
    out.update(rec)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return out


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        grid = spec["config"]["state_space_type"] == "grid"
        o = (run_grid_case if grid else run_case)(name, spec)
        print(f"{name}: lanes x steps = {o['done'].shape}, "
              f"done={int(o['done'].sum())}, resets={int(o['reset_after'].sum())}, "
              f"nonzero rewards={int((o['reward'] != 0).sum())}")
