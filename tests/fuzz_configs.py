"""Seeded random RLToyEnv configurations for the fuzz tests
(tests/test_fuzz_reference.py: reference vs oracle, build container;
tests/test_cuda_fuzz.py: CUDA vs oracle, GPU box).

The generators only emit what the reference documents as valid
(rl_toy_env.py:37-167); the seed lists below are the seeds whose configuration
the unmodified reference constructs and steps without raising (screened once
with tools/screen_fuzz_seeds.py in the build container -- the reference, not
this repo, decides what is a valid case).
"""
import numpy as np


def discrete_fuzz_config(seed):
    """Discrete env: sizes, diameter, sequence length, delay, both noises,
    reward cadence / density / distribution / scale / shift, terminal density
    and reward, connectivity, repeats, make_denser, irrelevant sub-MDP."""
    r = np.random.default_rng(1000 + seed)
    A = int(r.integers(3, 11))
    diam = int(r.choice([1, 1, 1, 2, 3]))
    tsd = float(r.choice([0.125, 0.25, 0.34, 0.5]))
    nonterm = A - int(tsd * A)
    L = int(r.integers(1, max(2, min(4, nonterm - 1)) + 1))
    cfg = dict(seed=int(r.integers(0, 1000)), state_space_type="discrete",
               action_space_type="discrete", action_space_size=A,
               state_space_size=A * diam, diameter=diam, sequence_length=L,
               delay=int(r.choice([0, 0, 1, 2, 3, 5])),
               terminal_state_density=tsd,
               reward_density=float(r.choice([0.1, 0.25, 0.5])),
               generate_random_mdp=True)
    if r.random() < 0.6:
        cfg["transition_noise"] = float(r.choice([0.0, 0.05, 0.2, 0.5]))
    if r.random() < 0.6:
        cfg["reward_noise"] = float(r.choice([0.0, 0.1, 1.0, 5.0]))
    if r.random() < 0.7:
        cfg["reward_every_n_steps"] = (True if r.random() < 0.4
                                       else int(r.integers(1, 5)))
    if r.random() < 0.4:
        cfg["reward_scale"] = float(r.choice([0.5, 2.5, -1.0]))
    if r.random() < 0.4:
        cfg["reward_shift"] = float(r.choice([-1.75, 0.25, 3.0]))
    if r.random() < 0.4:
        cfg["term_state_reward"] = float(r.choice([-3.0, 0.5, 10.0]))
    if r.random() < 0.3:
        cfg["make_denser"] = True
    if r.random() < 0.25 and diam == 1:
        cfg["repeats_in_sequences"] = True
    if r.random() < 0.25:
        cfg["maximally_connected"] = False
    if r.random() < 0.25:
        cfg["reward_dist"] = [float(r.choice([0.1, 0.5])), float(r.choice([1.0, 2.0]))]
    if r.random() < 0.25:  # irrelevant sub-MDP of its own size
        A1 = int(r.integers(2, 9))
        cfg["irrelevant_features"] = True
        cfg["action_space_size"] = [A, A1]
        cfg["state_space_size"] = [A * diam, A1 * diam]
    return cfg


def continuous_fuzz_config(seed):
    """Continuous env: dimension, relevant subset, dynamics order, inertia
    (scalar / per dimension), time unit, bounds, noises, delay, dense / sparse
    reward, action loss, terminal boxes, reward tail, move_along_a_line."""
    r = np.random.default_rng(5000 + seed)
    D = int(r.integers(1, 7))
    line = D >= 2 and r.random() < 0.3
    n_rel = int(r.integers(2 if line else 1, D + 1))
    rel = sorted(int(i) for i in r.choice(D, size=n_rel, replace=False))
    smax = float(r.choice([1.5, 3.0, 10.0]))
    cfg = dict(seed=int(r.integers(0, 1000)), state_space_type="continuous",
               action_space_type="continuous", state_space_dim=D,
               action_space_dim=D,
               transition_dynamics_order=int(r.integers(1, 4)),
               time_unit=float(r.choice([0.1, 0.25, 0.5, 1.0])),
               state_space_max=smax, action_space_max=float(r.choice([0.5, 1.0, 2.0])),
               delay=int(r.choice([0, 0, 1, 3])))
    if n_rel < D:
        cfg["relevant_indices"] = rel
        cfg["irrelevant_features"] = True
    else:
        n_rel, rel = D, list(range(D))
    k = r.random()
    if k < 0.3:
        cfg["inertia"] = float(r.choice([0.5, 2.0, 3.0]))
    elif k < 0.5:
        cfg["inertia"] = [float(x) for x in np.round(r.uniform(0.5, 3.0, size=D), 2)]
    else:
        cfg["inertia"] = 1.0
    if r.random() < 0.5:
        cfg["transition_noise"] = float(r.choice([0.01, 0.05, 0.2]))
    if r.random() < 0.5:
        cfg["reward_noise"] = float(r.choice([0.05, 0.5]))
    if r.random() < 0.4:
        cfg["reward_scale"] = float(r.choice([0.5, 2.0, -1.5]))
    if r.random() < 0.4:
        cfg["reward_shift"] = float(r.choice([-0.5, 0.25]))
    if r.random() < 0.3:
        cfg["term_state_reward"] = float(r.choice([-2.0, 5.0]))
    if r.random() < 0.3:
        n_box = int(r.integers(1, 3))
        cfg["terminal_states"] = [
            [float(x) for x in np.round(r.uniform(-0.8 * smax, 0.8 * smax, size=n_rel), 2)]
            for _ in range(n_box)]
        cfg["term_state_edge"] = float(r.choice([0.5, 1.0]))
    if line:
        cfg["reward_function"] = "move_along_a_line"
        cfg["sequence_length"] = int(r.integers(2, 8))
        if r.random() < 0.4:
            cfg["reward_every_n_steps"] = int(r.integers(1, 4))
    else:
        cfg["reward_function"] = "move_to_a_point"
        cfg["target_point"] = [float(x) for x in
                               np.round(r.uniform(-0.5 * smax, 0.5 * smax, size=n_rel), 2)]
        cfg["target_radius"] = float(r.choice([0.05, 0.5, 1.0]))
        if r.random() < 0.3:
            cfg["make_denser"] = False
        if r.random() < 0.3:
            cfg["action_loss_weight"] = float(r.choice([0.1, 0.5]))
    return cfg


def grid_fuzz_config(seed):
    """Grid env (rl_toy_env.py:1727-1778): shape, target, terminal cells,
    noises, dense / sparse reward, reward tail, irrelevant second grid."""
    r = np.random.default_rng(9000 + seed)
    shape = (int(r.integers(3, 10)), int(r.integers(3, 10)))
    target = [int(r.integers(0, shape[0])), int(r.integers(0, shape[1]))]
    cfg = dict(seed=int(r.integers(0, 1000)), state_space_type="grid",
               grid_shape=shape, delay=0, sequence_length=1,
               reward_function="move_to_a_point", target_point=target,
               make_denser=bool(r.random() < 0.6))
    if r.random() < 0.5:
        cfg["transition_noise"] = float(r.choice([0.1, 0.3, 0.6]))
    if r.random() < 0.5:
        cfg["reward_noise"] = float(r.choice([0.2, 1.0]))
    if r.random() < 0.4:
        cfg["reward_scale"] = float(r.choice([0.5, 3.0]))
    if r.random() < 0.4:
        cfg["reward_shift"] = float(r.choice([-0.5, 0.5]))
    if r.random() < 0.4:
        cfg["term_state_reward"] = float(r.choice([-0.25, 2.0]))
    if r.random() < 0.5:
        cells = {tuple(target)}
        for _ in range(int(r.integers(1, 4))):
            cells.add((int(r.integers(0, shape[0])), int(r.integers(0, shape[1]))))
        cfg["terminal_states"] = [list(c) for c in sorted(cells)]
    if r.random() < 0.3:
        cfg["reward_every_n_steps"] = int(r.integers(1, 4))
    if r.random() < 0.25:
        cfg["irrelevant_features"] = True
    return cfg


def image_fuzz_config(seed):
    """Image observations (rl_toy_env.py:705-717, :767-776): discrete envs with
    random transform subsets / quantisations / scale ranges / image sizes
    (also of both sub-states with irrelevant_features), continuous envs with
    target, terminal boxes and the irrelevant sub-image."""
    r = np.random.default_rng(13000 + seed)
    if r.random() < 0.7:
        A = int(r.integers(3, 9))
        cfg = dict(seed=int(r.integers(0, 1000)), state_space_type="discrete",
                   action_space_type="discrete", action_space_size=A,
                   state_space_size=A, sequence_length=1, delay=0,
                   terminal_state_density=0.25, reward_density=0.25,
                   generate_random_mdp=True, image_representations=True)
        tr = [t for t in ("shift", "scale", "rotate", "flip") if r.random() < 0.6]
        if tr:
            cfg["image_transforms"] = ",".join(tr)
        W, H = int(r.choice([64, 100, 128])), int(r.choice([64, 100, 120]))
        cfg["image_width"], cfg["image_height"] = W, H
        if r.random() < 0.7:
            cfg["image_sh_quant"] = int(r.choice([1, 2, 4, 8]))
        if r.random() < 0.7:
            cfg["image_ro_quant"] = int(r.choice([1, 5, 30, 90]))
        if r.random() < 0.7:
            lo = float(r.choice([0.5, 0.8, 1.0]))
            cfg["image_scale_range"] = (lo, float(lo + r.choice([0.0, 0.3, 0.5])))
        if r.random() < 0.3:
            cfg["transition_noise"] = 0.1
        if r.random() < 0.25:
            cfg["irrelevant_features"] = True
            cfg["action_space_size"] = [A, int(r.integers(2, 7))]
            cfg["state_space_size"] = list(cfg["action_space_size"])
        return cfg
    D = int(r.choice([2, 4]))
    smax = float(r.choice([2.0, 5.0]))
    cfg = dict(seed=int(r.integers(0, 1000)), state_space_type="continuous",
               action_space_type="continuous", state_space_dim=D, action_space_dim=D,
               transition_dynamics_order=int(r.integers(1, 3)), inertia=1.0,
               time_unit=float(r.choice([0.5, 1.0])), state_space_max=smax,
               action_space_max=1.0, reward_function="move_to_a_point",
               target_point=[float(x) for x in np.round(r.uniform(-smax, smax, size=2), 2)],
               target_radius=0.5, image_representations=True)
    if D == 4:
        cfg["relevant_indices"] = [0, 1]
        cfg["irrelevant_features"] = True
    if r.random() < 0.6:
        cfg["terminal_states"] = [
            [float(x) for x in np.round(r.uniform(-smax, smax, size=2), 2)]
            for _ in range(int(r.integers(1, 3)))]
        cfg["term_state_edge"] = float(r.choice([0.5, 1.5]))
    if r.random() < 0.5:
        cfg["image_width"], cfg["image_height"] = int(r.choice([64, 100])), int(r.choice([48, 100]))
    if r.random() < 0.3:
        cfg["transition_noise"] = 0.05
    return cfg


def wrapper_fuzz_spec(seed):
    """GymEnvWrapper tail (gym_env_wrapper.py:350-439, :523-618) around the
    deterministic stand-in base envs of tests/golden/make_wrapper_golden.py:
    discrete / Box / image bases, delay, action / observation noise, reward
    noise / scale / shift, terminal reward, padded image shift."""
    r = np.random.default_rng(17000 + seed)
    kind = str(r.choice(["discrete", "box", "image"]))
    cfg = dict(delay=int(r.choice([0, 1, 2, 4])))
    if r.random() < 0.6:
        cfg["reward_noise"] = float(r.choice([0.1, 0.5, 2.0]))
    if r.random() < 0.5:
        cfg["reward_scale"] = float(r.choice([0.5, 2.0, -1.0]))
    if r.random() < 0.5:
        cfg["reward_shift"] = float(r.choice([-1.0, 0.75]))
    if r.random() < 0.4:
        cfg["term_state_reward"] = float(r.choice([0.5, -2.0]))
    if kind == "box":
        cfg["state_space_type"] = "continuous"
        if r.random() < 0.7:
            cfg["transition_noise"] = float(r.choice([0.05, 0.1, 0.3]))
        return dict(base="box", dim=int(r.integers(1, 7)), config=cfg)
    cfg["state_space_type"] = "discrete"
    if r.random() < 0.6:
        cfg["transition_noise"] = float(r.choice([0.1, 0.2, 0.5]))
    spec = dict(base=kind, n_actions=int(r.integers(2, 9)), config=cfg)
    if kind == "image":
        spec["side"] = int(r.choice([8, 12, 16]))
        cfg["image_transforms"] = "shift"
        cfg["image_padding"] = int(r.choice([2, 5, 8]))
        cfg["image_sh_quant"] = int(r.choice([1, 2, 4]))
    return spec


# Screened with tools/screen_fuzz_seeds.py: the first seeds of each generator
# that the reference accepts and runs 40 steps of without raising (it rejects
# discrete seeds 3, 11, 14, 22, 27, 34, 40: too few rewardable sequences for
# the drawn density, an assertion of its own).
DISCRETE_SEEDS = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 16, 17, 18, 19, 20, 21, 23, 24, 25, 26, 28, 29, 30, 31, 32, 33, 35, 36, 37, 38, 39, 41, 42, 43, 44, 45, 46]
CONTINUOUS_SEEDS = list(range(40))
GRID_SEEDS = list(range(24))
IMAGE_SEEDS = list(range(24))
WRAPPER_SEEDS = list(range(16))
