"""Golden vectors of the reference GymEnvWrapper's post-processing tail
(envs/gym_env_wrapper.py:350-439, :523-618): action-substitution noise,
observation noise, reward delay FIFO, reward noise / scale / shift and the
padded image shift, recorded from the UNMODIFIED reference wrapped around small
deterministic stand-in environments (the wrapper is meant for external envs;
what the base env computes is irrelevant to the tail).

    python tests/golden/make_wrapper_golden.py

At HEAD the wrapper raises TypeError on every terminal step (`:414`
multiplies the reward_buffer LIST by a float); those steps are recorded as
`raised` with the inputs the tail saw, the episode is reset and the recording
continues.  Output: tests/golden/wrap_<case>.npz.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import draw_recorder  # noqa: E402
from oracle.ref_loader import load_reference_wrapper  # noqa: E402

LANE_SEED = 977

WRAPPER_CASES = {
    "wrap_discrete": dict(
        base="discrete", n_actions=6,
        config=dict(state_space_type="discrete", delay=2, transition_noise=0.2,
                    reward_noise=0.5, reward_scale=2.0, reward_shift=-1.0,
                    term_state_reward=0.5)),
    "wrap_discrete_plain": dict(
        base="discrete", n_actions=4,
        config=dict(state_space_type="discrete", delay=0, reward_scale=0.5)),
    "wrap_continuous": dict(
        base="box", dim=3,
        config=dict(state_space_type="continuous", delay=1, transition_noise=0.1,
                    reward_noise=0.25, reward_shift=0.75)),
    "wrap_image_shift": dict(
        base="image", n_actions=4, side=12,
        config=dict(state_space_type="discrete", delay=3, reward_noise=0.1,
                    image_transforms="shift", image_padding=5, image_sh_quant=2)),
}


def make_base(spec, seed):
    """Deterministic stand-in base envs (NOT reference code): observations,
    rewards and episode ends are simple functions of (seed, t, action)."""
    import gymnasium as gym
    from gymnasium.spaces import Box, Discrete

    class Base(gym.Env):
        def __init__(self):
            self.t, self.last_action = 0, None
            self.rng = np.random.default_rng(seed)
            if spec["base"] == "box":
                d = spec["dim"]
                self.action_space = Box(-np.ones(d, dtype=np.float32),
                                        np.ones(d, dtype=np.float32), dtype=np.float32)
                self.observation_space = Box(-np.full(d, 9, dtype=np.float32),
                                             np.full(d, 9, dtype=np.float32),
                                             dtype=np.float32)
            else:
                self.action_space = Discrete(spec["n_actions"])
                if spec["base"] == "image":
                    s = spec["side"]
                    self.observation_space = Box(np.zeros((s, s, 3), dtype=np.uint8),
                                                 np.full((s, s, 3), 255, dtype=np.uint8),
                                                 dtype=np.uint8)
                else:
                    self.observation_space = Discrete(50)

        def _obs(self):
            if spec["base"] == "box":
                return self.rng.uniform(-3, 3, size=spec["dim"]).astype(np.float32)
            if spec["base"] == "image":
                s = spec["side"]
                return self.rng.integers(0, 256, size=(s, s, 3)).astype(np.uint8)
            return int(self.rng.integers(50))

        def reset(self, seed=None, options=None):
            self.t = 0
            return self._obs(), {}

        def step(self, action):
            self.t += 1
            self.last_action = np.array(action).copy()
            reward = float(np.round(self.rng.normal(), 3))
            done = bool(self.rng.random() < 0.08)
            return self._obs(), reward, done, False, {}
    return Base()


def run_wrapper_case(name, spec, out_dir=HERE, lanes=4, steps=60):
    Wrapper = load_reference_wrapper()
    cfg = spec["config"]
    cont = cfg["state_space_type"] == "continuous"
    image = "image_transforms" in cfg
    rec = dict(action=[], applied=[], base_reward=[], base_done=[], base_obs=[],
               out_obs=[], out_reward=[], raised=[], choice_u=[], reward_noise=[],
               obs_noise=[], shift=[], reset_obs=[], reset_shift=[])
    for k in range(lanes):
        base = make_base(spec, seed=1000 + k)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            w = Wrapper(base, seed=LANE_SEED + k, **dict(cfg))
        log = []
        w._np_random = draw_recorder.RecordingGenerator(w._np_random, log, "w")
        arng = np.random.default_rng(50 + k)
        lane = {key: [] for key in rec}

        def do_reset():
            del log[:]
            with contextlib.redirect_stdout(io.StringIO()):
                obs, _ = w.reset()
            ints = [int(e[3]) for e in log if e[1] == "integers"]
            return obs, (ints + [0, 0])[:2]
        obs0, sh0 = do_reset()
        lane["reset_obs"].append(np.asarray(obs0))
        lane["reset_shift"].append(sh0)
        for t in range(steps):
            del log[:]
            a = arng.uniform(-1, 1, size=spec["dim"]).astype(np.float32) if cont \
                else int(arng.integers(spec["n_actions"]))
            # what the base env will answer, peeked through a copy of its state
            raised = False
            try:
                obs, r, done, trunc, _ = w.step(a)
            except TypeError:
                raised = True
                obs, r, done = None, np.nan, True
            lane["action"].append(np.asarray(a))
            lane["applied"].append(np.asarray(base.last_action))
            lane["raised"].append(raised)
            lane["base_done"].append(bool(done))
            us = [e[2] for e in log if e[1] == "choice_u"]
            lane["choice_u"].append(us[0] if us else np.nan)
            normals = [e[3] for e in log if e[1] == "normal"]
            lane["reward_noise"].append(
                next((float(x) for x in normals if np.ndim(x) == 0), np.nan))
            lane["obs_noise"].append(
                next((np.asarray(x) for x in normals if np.ndim(x) > 0),
                     np.full(spec.get("dim", 1), np.nan)))
            ints = [int(e[3]) for e in log if e[1] == "integers"]
            lane["shift"].append((ints + [0, 0])[:2])
            lane["out_reward"].append(float(r))
            lane["out_obs"].append(None if obs is None else np.asarray(obs))
            if raised:
                obs_r, sh_r = do_reset()
        for key in rec:
            rec[key].append(lane[key])
    # the base env's own outputs, replayed from the same seeds (the tail's inputs)
    for k in range(lanes):
        base = make_base(spec, seed=1000 + k)
        base.reset()
        br, bo = [], []
        for t in range(steps):
            o, r, d, _, _ = base.step(rec["applied"][k][t])
            br.append(r)
            bo.append(np.asarray(o))
            assert d == rec["base_done"][k][t]
            if d:
                base.reset()
        rec["base_reward"][k] = br
        rec["base_obs"][k] = bo
    out = {"config_delay": np.array(cfg.get("delay", 0))}
    for key in ("action", "applied", "base_reward", "base_done", "base_obs", "raised",
                "choice_u", "reward_noise", "obs_noise", "shift", "out_reward"):
        out[key] = np.array(rec[key])
    shape = np.asarray(next(o for o in rec["out_obs"][0] if o is not None)).shape
    oo = np.zeros((lanes, steps) + shape,
                  dtype=np.asarray(rec["out_obs"][0][0]).dtype
                  if rec["out_obs"][0][0] is not None else np.float64)
    for k in range(lanes):
        for t in range(steps):
            if rec["out_obs"][k][t] is not None:
                oo[k, t] = rec["out_obs"][k][t]
    out["out_obs"] = oo
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return out


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    for name, spec in WRAPPER_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        o = run_wrapper_case(name, spec)
        print(f"{name}: {o['raised'].shape}, terminal steps (reference raises) = "
              f"{int(o['raised'].sum())}, obs {o['out_obs'].shape}")
