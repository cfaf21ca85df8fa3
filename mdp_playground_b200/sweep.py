"""Experiment-grid front end: run a reference experiment
file's environment grid as ONE heterogeneous VectorRLToyEnv.

The reference expands `var_env_configs` of an `experiments/*.py` module into
a Cartesian grid (config_processor.get_grid_of_configs,
config_processor/config_processor.py:492-517), merges every cell over the
static `env_config["env_config"]` and hands each to a separate Ray trial
(scripts/run_experiments.py:231-276).  Here every cell becomes one
configuration group of a single batched env, stepped by one kernel launch,
and the per-cell episode statistics are written in the reference's CSV layout
(config_processor.py:241-259, :340-373) so its analysis code can read them.
Continuous and grid experiment files whose cells cannot share one launch
(different state dimension / dtype) get one batched env per cell instead.  No
agent is trained: the policy is uniform random (or supplied by the caller).
"""
import copy
import importlib.util
import itertools
import os
import sys
import types
from collections import OrderedDict

import numpy as np


def _ray_stub():
    """The experiment files do `from ray import tune` and only use
    `tune.grid_search([...])` in agent configs; ray itself is not needed to
    read the environment grid."""
    ray = types.ModuleType("ray")
    tune = types.ModuleType("ray.tune")
    tune.grid_search = lambda values: {"grid_search": list(values)}
    tune.choice = lambda values: {"choice": list(values)}
    tune.uniform = lambda a, b: {"uniform": (a, b)}
    tune.function = lambda f: f  # old Ray API the eval configs still use
    ray.tune = tune
    return {"ray": ray, "ray.tune": tune}


def load_experiment(path):
    """Import an experiment module (stubbing `ray` if it is not installed)."""
    path = os.path.abspath(path)
    if not path.endswith(".py"):
        path += ".py"
    stubs = {}
    try:
        import ray  # noqa: F401
    except ModuleNotFoundError:
        stubs = _ray_stub()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(
            "mdpp_experiment_" + os.path.basename(path)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def expand_grid(var_configs):
    """get_grid_of_configs (:492-517): Cartesian product of every leaf list,
    in the dict's iteration order.  Returns (keys, list of value tuples) with
    keys as (config_type, key) pairs."""
    keys, values = [], []
    for config_type, config_dict in var_configs.items():
        for key, leaf in config_dict.items():
            assert isinstance(leaf, list), (
                "var_configs should be a dict of dicts with lists as the leaf "
                "values to allow each configuration option to take multiple "
                "possible values")
            keys.append((config_type, key))
            values.append(leaf)
    cells = list(itertools.product(*values)) if values else []
    return keys, cells


def env_grid(module):
    """(var env keys, list of per-cell env config dicts) of an experiment."""
    var_configs = getattr(module, "var_configs", None)
    if var_configs is None:
        var_configs = OrderedDict({"env": module.var_env_configs})
    keys, cells = expand_grid(var_configs)
    static = copy.deepcopy(module.env_config.get("env_config", {}))
    env_keys = [k for t, k in keys if t == "env"]
    out = []
    for cell in cells:
        cfg = copy.deepcopy(static)
        for (t, k), v in zip(keys, cell):
            if t == "env":
                cfg[k] = v
        out.append(cfg)
    return env_keys, out


class Sweep:
    """All cells of an experiment's env grid in one VectorRLToyEnv."""

    def __init__(self, experiment, envs_per_cell=64, horizon=None, device=None,
                 shard=(0, 1), **env_kwargs):
        from .vector_env import VectorRLToyEnv
        self.module = load_experiment(experiment) if isinstance(experiment, str) \
            else experiment
        self.env_keys, self.cell_configs = env_grid(self.module)
        self.n_cells = len(self.cell_configs)
        if horizon is None:
            horizon = int(self.module.env_config.get("horizon", 100))
        self.envs_per_cell = int(envs_per_cell)
        kinds = {str(c.get("state_space_type", "")).lower()
                 for c in self.cell_configs}
        assert len(kinds) == 1, "one state_space_type per experiment file"
        self.kind = kinds.pop()
        if self.kind == "discrete":
            # the whole grid is ONE heterogeneous env: one launch per rollout
            self.env = VectorRLToyEnv(
                self.n_cells * self.envs_per_cell, device=device, autoreset=True,
                horizon=horizon, config_groups=self.cell_configs,
                group_sizes=[self.envs_per_cell] * self.n_cells, shard=shard,
                **env_kwargs)
            self.envs = [self.env]
        elif self._try_grouped(VectorRLToyEnv, device, horizon, shard, env_kwargs):
            pass  # the whole grid is ONE heterogeneous env here too
        else:
            # cells that differ in what shapes the state arrays (continuous:
            # dimension or dtype; grid: irrelevant_features) take one
            # configuration per context: one env per cell, launched back to
            # back on the same stream
            rank, world = shard
            self.envs = [VectorRLToyEnv(
                self.envs_per_cell, device=device, autoreset=True,
                horizon=horizon,
                env_id_offset=(i * world + rank) * self.envs_per_cell,
                **env_kwargs, **copy.deepcopy(c))
                for i, c in enumerate(self.cell_configs)]
            self.env = self.envs[0]
        self.timesteps = 0
        self._returned = np.zeros(self.n_cells)

    def _try_grouped(self, Env, device, horizon, shard, env_kwargs):
        """Continuous cells that share dim / dtype / reward function, and grid
        cells that share the number of coordinates, run as config groups of
        one env: one launch per rollout for the whole grid
        (e.g. experiments/sac_move_to_a_point_p_order_2.py:10-30)."""
        try:
            self.env = Env(
                self.n_cells * self.envs_per_cell, device=device, autoreset=True,
                horizon=horizon, config_groups=copy.deepcopy(self.cell_configs),
                group_sizes=[self.envs_per_cell] * self.n_cells, shard=shard,
                **env_kwargs)
        except (ValueError, NotImplementedError):
            return False
        self.envs = [self.env]
        self.grouped = True
        if self.kind == "grid":
            return True
        import torch
        amax = torch.tensor([float(c.get("action_space_max", np.inf))
                             for c in self.cell_configs], device=self.env.device)
        self._amax = amax.repeat_interleave(self.envs_per_cell).to(self.env._real)
        return True

    def run(self, steps_per_env, chunk=256, actions_fn=None):
        """Random-policy (or `actions_fn(t0, T) -> int32[T, N]`) rollouts."""
        import torch
        done = 0
        while done < steps_per_env:
            T = min(chunk, steps_per_env - done)
            if self.kind == "discrete" or getattr(self, "grouped", False):
                acts = None if actions_fn is None else actions_fn(done, T)
                if acts is None and self.kind == "continuous":
                    acts = self._random_actions_grouped(T, done)
                elif acts is None and self.kind == "grid":
                    acts = self._random_actions(self.env, T, done)
                out = self.env.rollout(T, actions=acts, want_final_obs=False)
                r = out["reward"].sum(dim=0).reshape(self.n_cells,
                                                     self.envs_per_cell)
                self._returned += r.sum(dim=1).to(torch.float64).cpu().numpy()
            else:
                for i, env in enumerate(self.envs):
                    acts = self._random_actions(env, T, done) \
                        if actions_fn is None else actions_fn(done, T, i)
                    out = env.rollout(T, actions=acts, want_final_obs=False)
                    self._returned[i] += float(out["reward"].sum())
            done += T
        self.timesteps += steps_per_env
        return self.results()

    def _random_actions(self, env, T, t0):
        """Uniform random policy for the per-cell envs (continuous: uniform
        in the action box, or N(0, 1) when it is unbounded; grid: one unit
        move or the no-op), seeded per env and chunk."""
        import torch
        N, dev = env.num_envs, env.device
        gen = torch.Generator(dev).manual_seed(
            (env.philox_seed + 7919 * env.env_id_offset + t0) % (2**63))
        if self.kind == "grid":
            nd = env._nd
            acts = torch.zeros((T, N, nd), dtype=torch.int64, device=dev)
            dim = torch.randint(0, nd, (T, N, 1), device=dev, generator=gen)
            acts.scatter_(2, dim, torch.randint(-1, 2, (T, N, 1), device=dev,
                                                generator=gen))
            return acts
        D, amax = env.spec.state_space_dim, env.spec.action_space_max
        if np.isfinite(amax):
            u = torch.rand((T, N, D), device=dev, generator=gen, dtype=env._real)
            return (u * 2 - 1) * amax
        return torch.randn((T, N, D), device=dev, generator=gen, dtype=env._real)

    def _random_actions_grouped(self, T, t0):
        """Uniform in every cell's own action box (N(0, 1) where unbounded)."""
        import torch
        env = self.env
        N, D, dev = env.num_envs, env.spec.state_space_dim, env.device
        gen = torch.Generator(dev).manual_seed(
            (env.philox_seed + 7919 * (env._shard[0] + 1) + t0) % (2**63))
        u = torch.rand((T, N, D), device=dev, generator=gen, dtype=env._real) * 2 - 1
        z = torch.randn((T, N, D), device=dev, generator=gen, dtype=env._real)
        bounded = torch.isfinite(self._amax)[None, :, None]
        return torch.where(bounded, u * torch.nan_to_num(self._amax, posinf=1.0)[None, :, None], z)

    def results(self, reduce=False):
        if self.kind == "discrete" or getattr(self, "grouped", False):
            st = self.env.episode_stats(reduce=reduce)
        else:  # one single-group env per cell
            per = [e.episode_stats(reduce=reduce) for e in self.envs]
            st = {k: np.array([float(np.asarray(p[k]).reshape(-1)[0]) for p in per])
                  for k in ("episodes", "transitions", "noisy_transitions")}
        ep = np.maximum(st["episodes"], 1)
        returned = self._returned
        if reduce:  # the reward sums are rank-local like the counters
            import torch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dev = self.env.device if self.kind == "discrete" else self.envs[0].device
                use_cuda = dist.get_backend() == "nccl"
                t = torch.as_tensor(returned, dtype=torch.float64,
                                    device=dev if use_cuda else "cpu").clone()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                returned = t.cpu().numpy()
        return {"episodes": st["episodes"], "transitions": st["transitions"],
                "episode_reward_mean": returned / ep,
                "episode_len_mean": st["transitions"] / ep,
                "noisy_transitions": st["noisy_transitions"]}

    def write_csv(self, path, algorithm="RandomPolicy"):
        """One row per grid cell in the reference's stats-file layout."""
        res = self.results()
        with open(path, "w") as f:
            f.write("# training_iteration, algorithm, "
                    + "".join(k + ", " for k in self.env_keys)
                    + "timesteps_total, episode_reward_mean, episode_len_mean\n")
            for i, cfg in enumerate(self.cell_configs):
                row = ["1", algorithm]
                for k in self.env_keys:
                    v = cfg[k]
                    row.append("%.2e" % v if isinstance(v, float)
                               else str(v).replace(" ", ""))
                row += [str(int(res["transitions"][i])),
                        "%.2e" % res["episode_reward_mean"][i],
                        "%.2e" % res["episode_len_mean"][i]]
                f.write(" ".join(row) + "\n")
        return path
