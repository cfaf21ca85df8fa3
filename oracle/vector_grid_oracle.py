"""Batched numpy oracle for grid environments (TEST INFRASTRUCTURE ONLY).

N environments sharing the configuration of one `ScalarRLToyEnv` (kind
"grid", itself pinned to the reference), with the batched-API semantics of the
CUDA path: same-step auto-reset, `horizon` truncation, noise from replayed
arrays or from the Philox streams (csrc/grid.cu: a Philox call per env and 2
steps for the noise decision + closed-form substitute action, one per 4 steps
for the reward normals).  Per-env arithmetic follows rl_toy_env.py:1727-1778
(transition), :1947-1965 + :1968-1990 (reward), :2098-2109 (done),
:2325-2345 (reset).
"""
import numpy as np

from . import philox as px

STREAM_GRID_STEP, STREAM_GRID_AUTORESET, STREAM_GRID_RESET = 40, 41, 42
STREAM_GRID_NORMAL, STREAM_GRID_ZIG = 43, 44


def substitute_action(w, a):
    """csrc/grid.cu substitute_action: uniform over the GridActionSpace
    samples (dim, value) that differ from the valid action `a`."""
    nd = len(a)
    nz = [k for k in range(nd) if a[k] != 0]
    if not nz:
        q = (int(w) * (2 * nd)) >> 32
        pick = (q >> 1) * 3 + (2 if q & 1 else 0)
    else:
        cur = nz[-1] * 3 + int(a[nz[-1]]) + 1
        q = (int(w) * (3 * nd - 1)) >> 32
        pick = q + (q >= cur)
    out = [0] * nd
    out[pick // 3] = pick % 3 - 1
    return out


class VectorGridOracle:
    def __init__(self, scalar_env, num_envs, autoreset=False, horizon=0, seed=0,
                 env_id_offset=0, fast_normal=False, normal="boxmuller"):
        e = scalar_env
        assert e.kind == "grid"
        # reward normals: "boxmuller" (fp64, what normal_precision="fp64" runs
        # in this kernel), "ziggurat" or, with fast_normal, the fp32 SFU
        # restatement
        self.normal = "fast" if fast_normal else normal
        assert self.normal in ("ziggurat", "boxmuller", "fast")
        self.N = int(num_envs)
        self.shape = [int(n) for n in e.grid_shape]
        self.nd = len(self.shape)
        self.target = [int(v) for v in e.target_point]
        self.dense = bool(e.make_denser)
        self.every_n = int(e.reward_every_n_steps)
        self.p = float(e.transition_noise) if e.transition_noise else 0.0
        self.has_pnoise = bool(e.transition_noise)
        self.pn_T = min(int(np.floor(self.p * 4294967296.0 + 0.5)), 4294967295)
        self.has_rnoise = e.has_reward_noise and e.reward_noise_std is not None
        self.r_std = e.reward_noise_std
        self.scale, self.shift = e.reward_scale, e.reward_shift
        self.term_reward = e.term_state_reward
        self.autoreset, self.horizon = autoreset, int(horizon)
        self.seed = int(seed)
        self.fast_normal = fast_normal
        self.gid = (np.arange(self.N, dtype=np.int64) + env_id_offset).astype(np.uint32)
        self.step_index = 0
        self.pos = np.zeros((self.N, self.nd), dtype=np.int64)
        self.t = np.zeros(self.N, dtype=np.int64)
        self.episode = np.zeros(self.N, dtype=np.int64)
        self.reached = np.zeros(self.N, dtype=bool)
        self.stats = dict(episodes=0, transitions=0, reward=0.0,
                          noisy_transitions=0, abs_reward_noise=0.0, terminated=0)

    def _cells(self, words):
        return np.stack([px.mulhi32(words[k], self.shape[k] + 1)
                         for k in range(self.nd)], axis=-1)

    def reset(self, mask=None, init_state=None):
        idx = np.arange(self.N) if mask is None else np.nonzero(mask)[0]
        if init_state is not None:
            s0 = np.asarray(init_state, dtype=np.int64)[idx]
        else:
            w = px.philox4x32_10(self.gid[idx], self.episode[idx].astype(np.uint32),
                                 0, STREAM_GRID_RESET, self.seed)
            s0 = self._cells(w)
        self.stats["episodes"] += int((self.t[idx] > 0).sum())
        self.pos[idx] = s0
        self.t[idx] = 0
        self.episode[idx] += 1
        self.reached[idx] = False
        return self.pos.copy()

    def rollout(self, T, actions, replay=None):
        N, nd = self.N, self.nd
        obs = np.zeros((T, N, nd), dtype=np.int64)
        final_obs = np.zeros((T, N, nd), dtype=np.int64)
        reward = np.zeros((T, N))
        term = np.zeros((T, N), dtype=bool)
        trunc = np.zeros((T, N), dtype=bool)
        tgt = np.array(self.target)
        for t in range(T):
            step = self.step_index + t
            # draws shared by 2 (noise decision, substitute) / 4 (normals) steps
            wp = px.step_words(self.seed, self.gid, step >> 1, STREAM_GRID_STEP)
            w = (wp[2], wp[3]) if step & 1 else (wp[0], wp[1])
            z0 = None
            if self.has_rnoise and self.normal == "ziggurat":
                wz = px.step_words(self.seed, self.gid, step >> 1, STREAM_GRID_ZIG)
                lo, hi = (wz[2], wz[3]) if step & 1 else (wz[0], wz[1])
                z0 = px.ziggurat_draw(self.seed, self.gid, lo, hi, step, 17)
            elif self.has_rnoise:
                wn = px.step_words(self.seed, self.gid, step >> 2, STREAM_GRID_NORMAL)
                pair = px.normal_pair_fast if self.fast_normal else px.normal_pair_f64
                j = step & 3
                z0 = pair(wn[0], wn[1])[j] if j < 2 else pair(wn[2], wn[3])[j - 2]
            wr = None
            for i in range(N):
                a = [int(v) for v in actions[t][i]]
                valid = all(-1 <= v <= 1 for v in a) and sum(abs(v) for v in a) <= 1
                if valid and self.has_pnoise:
                    if replay is not None:
                        if float(replay["noise_u"][t][i]) < self.p:
                            a = [int(v) for v in replay["noise_action"][t][i]]
                            self.stats["noisy_transitions"] += 1
                    elif int(w[0][i]) < self.pn_T:
                        a = substitute_action(w[1][i], a)
                        self.stats["noisy_transitions"] += 1
                d_old = int(np.abs(self.pos[i, :2] - tgt).sum())
                if valid:
                    for k in range(nd):
                        v = max(int(self.pos[i, k]) + a[k], 0)
                        self.pos[i, k] = min(v, self.shape[k] - 1)
                at = bool((self.pos[i, :2] == tgt).all())
                self.reached[i] |= at
                self.t[i] += 1
                r = 0.0
                if self.dense:
                    r = float(d_old - int(np.abs(self.pos[i, :2] - tgt).sum()))
                elif at:
                    r = 1.0
                if self.t[i] % self.every_n != 0:
                    r = 0.0
                self.stats["reward"] += r
                if self.has_rnoise:
                    n = float(replay["reward_noise"][t][i]) if replay is not None \
                        else self.r_std * float(z0[i])
                    self.stats["abs_reward_noise"] += abs(n)
                    r = r + n
                r = r * self.scale
                r = r + self.shift
                done = bool(self.reached[i])
                if done:
                    r = r + self.term_reward * self.scale
                tr = self.horizon > 0 and self.t[i] >= self.horizon
                reward[t, i], term[t, i], trunc[t, i] = r, done, tr
                final_obs[t, i] = self.pos[i]
                self.stats["terminated"] += int(done)
                self.stats["transitions"] += 1
                if self.autoreset and (done or tr):
                    if replay is not None:
                        self.pos[i] = replay["reset_state"][t][i]
                    else:
                        if wr is None:
                            wr = px.step_words(self.seed, self.gid, step,
                                               STREAM_GRID_AUTORESET)
                        self.pos[i] = self._cells([x[i:i + 1] for x in wr])[0]
                    self.t[i] = 0
                    self.episode[i] += 1
                    self.reached[i] = False
                    self.stats["episodes"] += 1
                obs[t, i] = self.pos[i]
        self.step_index += T
        return dict(obs=obs, final_obs=final_obs, reward=reward,
                    terminated=term, truncated=trunc)


class VectorGroupedGridOracle:
    """A heterogeneous grid launch (mdpp_set_grid_groups): group g = `sizes[g]`
    envs under the configuration of `scalar_envs[g]`, env-major; global Philox
    ids like mdp_playground_b200/sharding.py.  One VectorGridOracle per group."""

    def __init__(self, scalar_envs, sizes, autoreset=False, horizon=0, seed=0,
                 env_id_offset=0, rank=0, world=1, **kw):
        self.sizes = [int(n) for n in sizes]
        self.parts, gbegin = [], 0
        for e, n in zip(scalar_envs, self.sizes):
            self.parts.append(VectorGridOracle(
                e, n, autoreset=autoreset, horizon=horizon, seed=seed,
                env_id_offset=env_id_offset + gbegin + rank * n, **kw))
            gbegin += world * n

    def reset(self):
        return np.concatenate([p.reset() for p in self.parts])

    def rollout(self, T, actions):
        outs, begin = [], 0
        for p, n in zip(self.parts, self.sizes):
            outs.append(p.rollout(T, np.asarray(actions)[:, begin:begin + n]))
            begin += n
        return {k: np.concatenate([o[k] for o in outs], axis=1) for k in outs[0]}
