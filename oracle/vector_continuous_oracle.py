"""Batched numpy oracle for the continuous move_to_a_point step path
(TEST INFRASTRUCTURE ONLY).

Vectorised over N environments (elementwise numpy, Python loop over time
only), configured from a `ScalarRLToyEnv` (the scalar oracle, pinned to the
reference).  Arithmetic follows rl_toy_env.py:1625-1725 (transition),
:1912-1945 + :1968-1990 (reward), :2098-2109 (epilogue), :2284-2323 (reset)
in the reference's dtype path; on top of it the batched-API semantics of the
CUDA path: same-step auto-reset, horizon truncation, replayed or Philox noise
(oracle/philox.py restates the device streams).
"""
import math

import numpy as np

from . import philox as px


class VectorContinuousOracle:
    def __init__(self, scalar_env, num_envs, autoreset=False, horizon=0,
                 seed=0, env_id_offset=0, fast_normal=False, normal="boxmuller"):
        e = scalar_env
        assert e.kind == "continuous"
        # noise normals: "boxmuller" (fp64, what normal_precision="fp64" runs
        # in this kernel), "ziggurat" or, with fast_normal, the fp32 SFU
        # restatement
        self.normal = "fast" if fast_normal else normal
        assert self.normal in ("ziggurat", "boxmuller", "fast")
        self._pair = px.normal_pair_fast if fast_normal else px.normal_pair_f64
        self.e = e
        self.N, self.D = int(num_envs), e.state_space_dim
        self.R = np.dtype(e.dtype_s).type
        self.order = e.dynamics_order
        self.rel = list(e.relevant_indices)
        self.delay, self.every_n = e.delay, int(e.reward_every_n_steps)
        self.autoreset, self.horizon = autoreset, int(horizon)
        self.seed = int(seed)
        self.gid = (np.arange(self.N, dtype=np.int64) + env_id_offset).astype(
            np.uint32)
        self.step_index = 0
        self.line = e.config["reward_function"] == "move_along_a_line"
        self.L = int(e.sequence_length)
        self.target64 = "target_point" not in e.config and not self.line
        # last L emitted relevant states (move_along_a_line), slot = step % L
        self.hist = np.zeros((max(self.L, 1), self.N, len(self.rel)), dtype=self.R)
        self.sd = np.zeros((self.order + 1, self.N, self.D), dtype=self.R)
        self.em = np.zeros((self.N, self.D), dtype=self.R)
        self.t = np.zeros(self.N, dtype=np.int64)
        self.episode = np.zeros(self.N, dtype=np.int64)
        self.reached = np.zeros(self.N, dtype=bool)
        self.ring = np.zeros((max(self.delay, 1), self.N), dtype=self.R)
        self.ring64 = np.zeros((max(self.delay, 1), self.N), dtype=np.float64)
        self.has_pnoise = e.has_transition_noise and e.transition_noise is not None
        self.has_rnoise = e.has_reward_noise and e.reward_noise_std is not None

    # -- helpers -----------------------------------------------------------
    def _norm(self, x, dtype):
        """Sequential, separately rounded sum of squares (what numpy does for
        the 1-2 element vectors of this path), sqrt in the same dtype."""
        s = np.zeros(x.shape[0], dtype=dtype)
        for k in range(x.shape[1]):
            s = (s + (x[:, k] * x[:, k]).astype(dtype)).astype(dtype)
        return np.sqrt(s).astype(dtype)

    def _line_reward(self, step):
        """rl_toy_env.py:1865-1910 per env, with numpy's own float32 SVD like
        oracle/scalar_env.py: python-float (float64) rewards."""
        from .scalar_env import dist_of_pt_from_line
        L, out = self.L, np.zeros(self.N)
        order = [(step - L + 1 + k) % L for k in range(L)]   # oldest first
        for i in np.nonzero(self.t >= L)[0]:
            # (the reference's window is F-ordered -- a slice combined with an
            # index list -- so its mean runs numpy's pairwise summation)
            data = np.asfortranarray(self.hist[order, i, :])
            mean = data.mean(axis=0)
            _, _, vv = np.linalg.svd(data - mean)
            ends = vv[0] * np.linspace(-1, 1, 2)[:, np.newaxis]
            ends += mean
            total = 0
            for pt in data:
                total += dist_of_pt_from_line(pt, ends[0], ends[-1])
            out[i] = 0.0 + -total / L
        return out

    def _dist(self, x):
        R = self.R
        if self.line:
            return np.full(x.shape[0], np.inf)   # no target, nothing to reach
        if self.target64:
            diff = x[:, self.rel].astype(np.float64) - self.e.target_point
            return self._norm(diff, np.float64)
        diff = (x[:, self.rel] - self.e.target_point.astype(R)).astype(R)
        return self._norm(diff, R).astype(np.float64)

    def _in_box(self, x):
        any_ = np.zeros(x.shape[0], dtype=bool)
        for lo, hi in zip(self.e.term_lows, self.e.term_highs):
            v = x[:, self.rel]
            any_ |= np.all((v >= lo) & (v <= hi), axis=1)
        return any_

    def _box_sample(self, idx, attempt):
        R, D = self.R, self.D
        smax = self.e.state_space_max
        out = np.zeros((len(idx), D), dtype=R)
        hi = float(R(smax)) if np.isfinite(smax) else np.inf
        for c in range((D + 1) // 2):
            w = px.philox4x32_10(self.gid[idx], self.episode[idx].astype(np.uint32),
                                 np.uint32(attempt), px.STREAM_RESET_BOX + c,
                                 self.seed)
            if np.isfinite(smax):
                v0 = -hi + (hi + hi) * px.uniform53(w[0], w[1])
                v1 = -hi + (hi + hi) * px.uniform53(w[2], w[3])
            else:
                v0, v1 = px.normal_pair_f64(w[0], w[1])
            out[:, 2 * c] = v0.astype(R)
            if 2 * c + 1 < D:
                out[:, 2 * c + 1] = v1.astype(R)
        return out

    def _sample_nonterminal(self, idx):
        s0 = self._box_sample(idx, 0)
        if self.e.term_lows:
            for attempt in range(1, 64):
                bad = self._in_box(s0)
                if not bad.any():
                    break
                s0[bad] = self._box_sample(idx[bad], attempt)
        return s0

    def _do_reset(self, idx, s0):
        self.em[idx] = s0
        self.sd[:, idx, :] = 0
        self.sd[0, idx] = s0
        self.t[idx] = 0
        self.reached[idx] = False
        self.episode[idx] += 1

    def reset(self, mask=None, init_state=None):
        idx = np.arange(self.N) if mask is None else np.nonzero(mask)[0]
        s0 = (np.asarray(init_state, dtype=self.R)[idx] if init_state is not None
              else self._sample_nonterminal(idx))
        self._do_reset(idx, s0)
        return self.em.copy()

    # -- the step ----------------------------------------------------------
    def rollout(self, T, actions, replay=None):
        R, N, D, n = self.R, self.N, self.D, self.order
        e = self.e
        obs = np.zeros((T, N, D), dtype=R)
        final_obs = np.zeros((T, N, D), dtype=R)
        reward = np.zeros((T, N), dtype=R)
        term = np.zeros((T, N), dtype=bool)
        trunc = np.zeros((T, N), dtype=bool)
        amax, smax = R(e.action_space_max), R(e.state_space_max)
        radius = e.target_radius if self.target64 else float(R(e.target_radius))
        for t in range(T):
            step = self.step_index + t
            a = np.asarray(actions[t], dtype=R)
            in_range = np.all((a >= -amax) & (a <= amax), axis=1)
            dist_old = self._dist(self.em)
            sd = self.sd.copy()
            # a / inertia (:1654): scalar or dtype_s vector -> dtype_s division;
            # a list / float64 vector -> float64 top derivative (numpy promotion)
            top64 = None
            if np.ndim(e.inertia) == 0:
                sd[n] = (a / R(e.inertia)).astype(R)
            elif np.asarray(e.inertia).dtype == np.dtype(R):
                sd[n] = (a / np.asarray(e.inertia)[None, :]).astype(R)
            else:
                top64 = (a.astype(np.float64)
                         / np.asarray(e.inertia, dtype=np.float64)[None, :])
                sd[n] = top64.astype(R)
            for i in range(n):
                for j in range(n - i):
                    if top64 is not None and i + j + 1 == n:
                        term64 = top64 * (e.time_unit ** (j + 1))
                    else:
                        term64 = (sd[i + j + 1] * R(e.time_unit ** (j + 1))).astype(
                            R).astype(np.float64)
                    sd[i] = (sd[i].astype(np.float64)
                             + term64 / float(math.factorial(j + 1))).astype(R)
            self.sd[:, in_range] = sd[:, in_range]
            nxt = np.where(in_range[:, None], self.sd[0], self.em).astype(R)
            if self.has_pnoise:
                if replay is not None:
                    nz = np.asarray(replay["state_noise"][t], dtype=np.float64)
                else:
                    nz = np.zeros((N, D))
                    for c in range((D + 1) // 2 if self.normal == "ziggurat" else 0):
                        w = px.step_words(self.seed, self.gid, step,
                                          px.STREAM_STATE_NOISE + c)
                        nz[:, 2 * c] = e.transition_noise * px.ziggurat_draw(
                            self.seed, self.gid, w[0], w[1], step, 2 * c)
                        if 2 * c + 1 < D:
                            nz[:, 2 * c + 1] = e.transition_noise * px.ziggurat_draw(
                                self.seed, self.gid, w[2], w[3], step, 2 * c + 1)
                    for c in range(0 if self.normal == "ziggurat" else (D + 3) // 4):
                        w = px.step_words(self.seed, self.gid, step,
                                          px.STREAM_STATE_NOISE + c)
                        z01 = self._pair(w[0], w[1])
                        z23 = self._pair(w[2], w[3])
                        for k, z in enumerate((*z01, *z23)):
                            if 4 * c + k < D:
                                nz[:, 4 * c + k] = e.transition_noise * z
                nxt = (nxt.astype(np.float64) + nz).astype(R)
            in_bounds = np.all((nxt >= -smax) & (nxt <= smax), axis=1)
            if e.image_representations:
                in_bounds[:] = False
            out = ~in_bounds
            if out.any():
                clipped = np.clip(nxt[out], -smax, smax).astype(R)
                nxt[out] = clipped
                self.sd[:, out, :] = 0
                self.sd[0, out] = clipped
            dist_new = self._dist(nxt)
            self.reached |= dist_new < radius
            self.t += 1
            # reward, carried as (value, is_real) like numpy's scalar typing
            if self.line:
                self.hist[step % self.L] = nxt[:, self.rel]
                r64, is_real = self._line_reward(step), np.zeros(N, dtype=bool)
                rr = np.zeros(N, dtype=R)
            elif e.make_denser:
                if self.target64:
                    r64, is_real = -dist_new + dist_old, np.zeros(N, dtype=bool)
                    rr = np.zeros(N, dtype=R)
                else:
                    rr = (-dist_new.astype(R) + dist_old.astype(R)).astype(R)
                    r64, is_real = np.zeros(N), np.ones(N, dtype=bool)
            else:
                r64 = (dist_new < radius).astype(np.float64)
                rr, is_real = np.zeros(N, dtype=R), np.zeros(N, dtype=bool)
            loss = (R(e.action_loss_weight) * self._norm(a, R)).astype(R)
            if self.line:
                pass  # (the action loss belongs to move_to_a_point, :1943)
            elif e.make_denser and self.target64 and R is np.float32:
                r64 = r64 - loss.astype(np.float64)
            else:
                rr = np.where(is_real, rr, r64.astype(R))
                rr = (rr - loss).astype(R)
                is_real = np.ones(N, dtype=bool)
            if self.delay > 0:
                pos = step % self.delay
                have = self.t > self.delay
                if (self.line or (e.make_denser and self.target64)) and R is np.float32:
                    # python floats all the way through reward_buffer (:1968-1977)
                    delayed = self.ring64[pos].copy()
                    self.ring64[pos] = r64
                    r64 = np.where(have, delayed, 0.0)
                    is_real = np.zeros(N, dtype=bool)
                else:
                    delayed = self.ring[pos].copy()
                    self.ring[pos] = np.where(is_real, rr, r64.astype(R))
                    rr = np.where(have, delayed, rr)
                    r64 = np.where(have, r64, 0.0)
                    is_real = have.copy()
            gated = (self.t % self.every_n) != 0
            r64 = np.where(gated, 0.0, r64)
            is_real = is_real & ~gated
            if self.has_rnoise:
                if replay is not None:
                    nrw = np.asarray(replay["reward_noise"][t], dtype=np.float64)
                else:
                    w = px.step_words(self.seed, self.gid, step, px.STREAM_NORMAL)
                    if self.normal == "ziggurat":
                        nrw = e.reward_noise_std * px.ziggurat_draw(
                            self.seed, self.gid, w[0], w[1], step, 16)
                    else:
                        nrw = e.reward_noise_std * self._pair(w[0], w[1])[0]
            else:
                nrw = np.zeros(N)
            done = self._in_box(nxt) | self.reached
            term_add = e.term_state_reward * e.reward_scale
            # np.float32 path
            a32 = rr
            if self.has_rnoise:
                a32 = (a32 + nrw.astype(R)).astype(R)
            a32 = (a32 * R(e.reward_scale)).astype(R)
            a32 = (a32 + R(e.reward_shift)).astype(R)
            a32 = np.where(done, (a32 + R(term_add)).astype(R), a32)
            # python-float path
            a64 = r64
            if self.has_rnoise:
                a64 = a64 + nrw
            a64 = a64 * e.reward_scale
            a64 = a64 + e.reward_shift
            a64 = np.where(done, a64 + term_add, a64)
            reward[t] = np.where(is_real, a32, a64.astype(R))
            tr = (self.t >= self.horizon) if self.horizon > 0 \
                else np.zeros(N, dtype=bool)
            term[t], trunc[t] = done, tr
            final_obs[t] = nxt
            self.em = nxt.copy()
            if self.autoreset:
                idx = np.nonzero(done | tr)[0]
                if len(idx):
                    s0 = (np.asarray(replay["reset_state"][t], dtype=R)[idx]
                          if replay is not None else self._sample_nonterminal(idx))
                    self._do_reset(idx, s0)
            obs[t] = self.em
        self.step_index += T
        return dict(obs=obs, final_obs=final_obs, reward=reward,
                    terminated=term, truncated=trunc)


class VectorGroupedContinuousOracle:
    """A heterogeneous continuous launch: group g = `sizes[g]` envs under the
    configuration of `scalar_envs[g]`, env-major; global Philox ids like
    mdp_playground_b200/sharding.py.  One VectorContinuousOracle per group."""

    def __init__(self, scalar_envs, sizes, autoreset=False, horizon=0, seed=0,
                 env_id_offset=0, rank=0, world=1, **kw):
        self.sizes = [int(n) for n in sizes]
        self.parts, gbegin = [], 0
        for e, n in zip(scalar_envs, self.sizes):
            self.parts.append(VectorContinuousOracle(
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
