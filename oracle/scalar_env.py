"""Scalar CPU oracle: a restatement of the reference `RLToyEnv` step path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  One Python object = one
environment, exactly like the reference, consuming the same numpy PCG64
streams in the same order, so that with `draws="numpy"` it reproduces the
reference's trajectories bit for bit (states, fp64/fp32 rewards, image
pixels).  With `draws=<ReplayDraws>` the noise comes from recorded values,
with `draws=<PhiloxDraws>` from the counter-based streams the CUDA kernels
use (oracle/philox.py).

Reference files restated (all under /root/reference/mdp_playground):
  envs/rl_toy_env.py   __init__ :216-853, init_terminal_states :855-990,
                       init_init_state_dist :992-1040,
                       init_transition_function :1042-1248,
                       init_reward_function :1253-1573,
                       transition_function :1577-1725,
                       reward_function :1782-1990, step :1992-2125,
                       reset :2217-2377
  spaces/discrete_extended.py :11-23, spaces/image_multi_discrete.py :129-288,
  spaces/image_continuous.py :116-277
Third-party arithmetic restated: gymnasium 1.x `seeding.np_random`,
`Box.sample`/`Box.contains` (unpinned in the reference's setup.py:112);
Pillow (12.2.0 in this image) is called directly for polygon/ellipse/rotate,
exactly as the reference does.

Not covered (SURVEY.md section 8f "next" rows):
callable P/R/noise.
"""
import math
import sys

import numpy as np


def np_random(seed):
    """gymnasium.utils.seeding.np_random (rl_toy_env.py:2399)."""
    ss = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(ss)), ss.entropy


# --------------------------------------------------------------------------
# draw sources: every random number the step/reset path consumes
# --------------------------------------------------------------------------
class NumpyDraws:
    """The reference's own streams: E = env RNG,
    S = observation_spaces[0] RNG, I = image-space RNG, F = feature-space RNG."""

    def __init__(self, env):
        self.env = env

    def transition_uniform(self):
        # Generator.choice(p=...) draws exactly one random()
        return self.env.rng_S.random()

    def irr_transition_uniform(self):
        return self.env.rng_S1.random()

    def irr_reset_uniform(self):
        return self.env.rng_E.random()

    def reward_normal(self, std):
        return self.env.rng_E.normal(0, std)

    def state_normal(self, std, dim):
        return self.env.rng_E.normal(0, std, (dim,))

    def reset_uniform(self):
        return self.env.rng_E.random()

    def reset_box(self, env):
        return env._box_sample(self.env.rng_F)

    def grid_noise_uniform(self):  # rl_toy_env.py:1736 (E stream)
        return self.env.rng_E.uniform()

    def grid_action_sample(self, ndim):  # GridActionSpace.sample (A stream)
        ind = self.env.rng_A.integers(ndim).item()
        return ind, self.env.rng_A.integers(3).item() - 1

    def image_scale_u(self):
        return self.env.rng_I.random()

    def image_integers(self, low, high=None):
        return self.env.rng_I.integers(low, high).item()


class ReplayDraws:
    """Feeds recorded draws.  `feed` maps a draw name to a list consumed
    front to back."""

    def __init__(self, feed):
        self.feed = {k: list(v) for k, v in feed.items()}

    def _pop(self, k):
        return self.feed[k].pop(0)

    def transition_uniform(self):
        return self._pop("transition_u")

    def irr_transition_uniform(self):
        return self._pop("irr_transition_u")

    def irr_reset_uniform(self):
        return self._pop("irr_reset_u")

    def reward_normal(self, std):
        return self._pop("reward_noise")

    def state_normal(self, std, dim):
        return np.asarray(self._pop("state_noise"), dtype=np.float64)

    def reset_uniform(self):
        return self._pop("reset_u")

    def reset_box(self, env):
        return np.asarray(self._pop("reset_state"))

    def grid_noise_uniform(self):
        return self._pop("grid_noise_u")

    def grid_action_sample(self, ndim):
        return tuple(self._pop("grid_noise_action"))

    def image_scale_u(self):
        return self._pop("image_scale_u")

    def image_integers(self, low, high=None):
        return int(self._pop("image_int"))


def dist_of_pt_from_line(pt, ptA, ptB):
    """rl_toy_env.py:2546-2576: distance of `pt` from the line through ptA and
    ptB by Pythagoras (float64 here: the end points are float64)."""
    tolerance = 1e-13
    lineAB = ptA - ptB
    lineApt = ptA - pt
    dot_product = np.dot(lineAB, lineApt)
    if np.linalg.norm(lineAB) < tolerance:
        return 0
    proj = dot_product / np.linalg.norm(lineAB)
    sq_dist = np.linalg.norm(lineApt) ** 2 - proj ** 2
    if sq_dist < 0:
        sq_dist = 0
    return np.sqrt(sq_dist)


def choice_from_uniform(cdf, u):
    """numpy Generator.choice(p=...) given its one uniform draw
    (discrete_extended.py:17; SURVEY.md 8a row A1')."""
    return int(np.searchsorted(cdf, u, side="right"))


def normalised_cdf(p):
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


class ScalarRLToyEnv:
    def __init__(self, draws="numpy", **config):
        cfg = config
        if cfg == {}:  # rl_toy_env.py:227-235
            cfg = dict(state_space_size=8, action_space_size=8,
                       state_space_type="discrete",
                       action_space_type="discrete",
                       terminal_state_density=0.25, maximally_connected=True)
        self.config = cfg
        self._seed_fanout(cfg)  # :285-334
        cfg["state_space_type"] = cfg["state_space_type"].lower()
        kind = cfg["state_space_type"]
        self.kind = kind
        g = cfg.get
        self.use_custom_mdp = g("use_custom_mdp", False)
        if self.use_custom_mdp:
            assert "transition_function" in cfg and "reward_function" in cfg
        self.terminal_state_density = g("terminal_state_density", 0.25)
        self.term_state_reward = g("term_state_reward", 0.0)
        self.delay = g("delay", 0)
        self.sequence_length = g("sequence_length", 1)
        self.reward_density = g("reward_density", 0.25)
        if "make_denser" in cfg:
            self.make_denser = cfg["make_denser"]
        else:
            self.make_denser = (kind == "continuous")  # :385-391
        self.maximally_connected = g("maximally_connected", True)
        # q2: the *presence* of the key decides whether a draw is consumed
        self.has_reward_noise = "reward_noise" in cfg
        self.reward_noise_std = cfg.get("reward_noise")
        self.has_transition_noise = "transition_noise" in cfg
        self.transition_noise = cfg.get("transition_noise")
        assert not callable(self.reward_noise_std), "callable noise: next row"
        assert not callable(self.transition_noise), "callable noise: next row"
        self.reward_scale = g("reward_scale", 1.0)
        self.reward_shift = g("reward_shift", 0.0)
        self.irrelevant_features = bool(g("irrelevant_features", False))
        self.image_representations = g("image_representations", False)
        if "image_transforms" in cfg:
            assert kind == "discrete"
            self.image_transforms = cfg["image_transforms"]
        else:
            self.image_transforms = "none"
        self.image_width = g("image_width", 100)
        self.image_height = g("image_height", 100)
        if kind == "discrete":  # :461-508
            self.image_sh_quant = g(
                "image_sh_quant", 1 if "shift" in self.image_transforms else None)
            self.image_ro_quant = g(
                "image_ro_quant", 1 if "rotate" in self.image_transforms else None)
            self.image_scale_range = g(
                "image_scale_range",
                (0.5, 1.5) if "scale" in self.image_transforms else None)
            self.reward_dist = g("reward_dist", None)
            self.diameter = g("diameter", 1)
        elif kind == "continuous":
            self.state_space_dim = cfg["state_space_dim"]
            cfg.setdefault("reward_function", "move_to_a_point")
            assert cfg["reward_function"] in ("move_to_a_point", "move_along_a_line")
            self.dynamics_order = g("transition_dynamics_order", 1)
            self.inertia = g("inertia", 1.0)
            self.time_unit = g("time_unit", 1.0)
            self.target_radius = g("target_radius", 0.05)
        elif kind == "grid":  # :539-541, :655-657
            assert "grid_shape" in cfg
            self.grid_shape = tuple(cfg["grid_shape"])
            assert cfg["reward_function"] == "move_to_a_point"
            self.target_point = cfg["target_point"]
            # the reference leaves make_denser unset for grid envs (:384-390)
            # and crashes in the reward without it; delay / sequence_length
            # other than 0 / 1 crash in np.array(augmented_state) (:1950)
            assert "make_denser" in cfg and self.delay == 0 \
                and self.sequence_length == 1
        else:
            raise ValueError("Unknown state_space_type")
        self.action_loss_weight = g("action_loss_weight", 0.0)
        if "reward_every_n_steps" in cfg:
            self.reward_every_n_steps = cfg["reward_every_n_steps"]
        else:  # :550-561
            self.reward_every_n_steps = (
                self.sequence_length if kind == "discrete" else 1)
        self.repeats_in_sequences = g("repeats_in_sequences", False)

        if kind == "discrete":  # :570-591
            self.dtype_s = g("dtype_s", np.int64)
            if self.irrelevant_features:  # :574-578, sizes are [relevant, irrelevant]
                assert len(cfg["action_space_size"]) == 2
                assert not self.use_custom_mdp
                self.action_space_size = cfg["action_space_size"][0]
                self.action_space_size_irr = cfg["action_space_size"][1]
                self.state_space_size_irr = (self.action_space_size_irr
                                             * self.diameter)
            else:
                assert isinstance(cfg["action_space_size"], int)
                self.action_space_size = cfg["action_space_size"]
            if self.use_custom_mdp:
                self.state_space_size = cfg["state_space_size"]
            else:
                self.state_space_size = self.action_space_size * self.diameter
        elif kind == "grid":  # :604-608
            self.dtype_s = g("dtype_s", np.int64)
            if g("irrelevant_features", False):
                self.grid_shape = self.grid_shape * 2
        else:  # :593-602
            self.dtype_s = g("dtype_s", np.float32)
            if "relevant_indices" not in cfg:
                cfg["relevant_indices"] = range(self.state_space_dim)
            self.relevant_indices = list(cfg["relevant_indices"])
        self.dtype_o = g("dtype_o", np.uint8 if self.image_representations
                         else self.dtype_s)
        if "init_state_dist" in cfg and "relevant_init_state_dist" not in cfg:
            cfg["relevant_init_state_dist"] = cfg["init_state_dist"]
        assert self.sequence_length > 0
        if kind == "continuous" and cfg["reward_function"] == "move_to_a_point":
            assert self.sequence_length == 1  # :641-654
            if "target_point" in cfg:
                self.target_point = np.array(cfg["target_point"],
                                             dtype=self.dtype_s)
                assert self.target_point.shape == (len(self.relevant_indices),)
            else:
                self.target_point = np.zeros((self.state_space_dim,))
        self.augmented_state_length = self.sequence_length + self.delay + 1
        self.total_episodes = 0

        self._init_terminal_states()
        # spaces (:668-776): only their generators matter here
        if kind == "discrete":
            self.rng_S, _ = np_random(self.seed_dict["relevant_state_space"])
            if self.irrelevant_features:
                self.rng_S1, _ = np_random(self.seed_dict["irrelevant_state_space"])
                if not self.image_representations:
                    # :738-741 wraps both sub-spaces in a TupleExtended seeded
                    # with an int: gymnasium 1.x Tuple.seed then RE-SEEDS every
                    # sub-space with integers(int32.max, size=2) of its own
                    # stream (gymnasium_standin/spaces/tuple.py; parity of this
                    # cascade is unpinned by the reference)
                    rng_T, _ = np_random(self.seed_dict["state_space"])
                    sub = rng_T.integers(np.iinfo(np.int32).max, size=2)
                    self.rng_S, _ = np_random(int(sub[0]))
                    self.rng_S1, _ = np_random(int(sub[1]))
            if self.image_representations:
                self.rng_I, _ = np_random(self.seed_dict["image_representations"])
        elif kind == "grid":  # :780-812: Box(0, grid_shape) int64, GridActionSpace
            self.rng_F, _ = np_random(self.seed_dict["state_space"])
            self.rng_A, _ = np_random(self.seed_dict["action_space"])
        else:
            self.state_space_max = g("state_space_max", np.inf)
            self.action_space_max = g("action_space_max", np.inf)
            self.rng_F, _ = np_random(self.seed_dict["state_space"])
            D = self.state_space_dim
            self.feat_low = np.full((D,), -self.state_space_max).astype(self.dtype_s)
            self.feat_high = np.full((D,), self.state_space_max).astype(self.dtype_s)
            self.act_low = np.full((D,), -self.action_space_max).astype(self.dtype_s)
            self.act_high = np.full((D,), self.action_space_max).astype(self.dtype_s)
            if self.image_representations:
                assert (self.feat_high != np.inf).any()
                self._image_cont_setup()
        self.draws = NumpyDraws(self) if draws == "numpy" else draws
        self._init_init_state_dist()
        self._init_transition_function()
        self._init_reward_function()
        # :831-833 -- re-seeds E with seed_dict["env"], then first reset
        self.reset(seed=self.seed_dict["env"])

    # ---- seeding ----------------------------------------------------------
    def _seed_fanout(self, cfg):
        """rl_toy_env.py:285-334."""
        if "seed" in cfg and isinstance(cfg["seed"], dict):
            self.seed_dict = cfg["seed"]
            self.rng_E, self.seed_ = np_random(self.seed_dict["env"])
            return
        seed_int = cfg.get("seed")
        if seed_int is not None and not isinstance(seed_int, int):
            raise TypeError("Unsupported data type for seed", type(seed_int))
        self.seed_dict = {"env": seed_int}
        self.rng_E, self.seed_ = np_random(seed_int)
        for k in ("relevant_state_space", "relevant_action_space",
                  "irrelevant_state_space", "irrelevant_action_space",
                  "state_space", "action_space", "image_representations"):
            self.seed_dict[k] = self.rng_E.integers(sys.maxsize).item()

    # ---- tables -----------------------------------------------------------
    def _init_terminal_states(self):
        """rl_toy_env.py:855-956."""
        cfg = self.config
        if self.kind == "discrete":
            if self.use_custom_mdp and "terminal_states" in cfg:
                ts = cfg["terminal_states"]
                assert not callable(ts), "callable terminal_states: next row"
                self.terminal_states = np.array(ts)
            else:
                A = self.action_space_size
                self.num_terminal_states = int(self.terminal_state_density * A)
                self.terminal_states = np.array(
                    [j * A - 1 - i for j in range(1, self.diameter + 1)
                     for i in range(self.num_terminal_states)])
                cfg["terminal_states"] = self.terminal_states
            self.is_terminal_state = lambda s: s in self.terminal_states
        elif self.kind == "grid":
            # :958-987: the terminal cells become int64 Boxes, but the state is
            # tested as a FLOAT array, which Box.contains rejects (can_cast
            # float64 -> int64 is False): no grid state is ever terminal.  The
            # cells still show up in the image observations.
            assert not callable(cfg.get("terminal_states"))
            self.term_cells = [list(c) for c in cfg.get("terminal_states", [])]
            self.is_terminal_state = lambda s: False
        else:
            self.term_lows, self.term_highs = [], []
            if "terminal_states" in cfg:
                assert not callable(cfg["terminal_states"])
                edge = cfg["term_state_edge"]
                for centre in cfg["terminal_states"]:
                    assert len(centre) == len(self.relevant_indices)
                    lo = np.array([c - edge / 2 for c in centre])
                    hi = np.array([c + edge / 2 for c in centre])
                    # Box casts its bounds to dtype_s
                    self.term_lows.append(lo.astype(self.dtype_s))
                    self.term_highs.append(hi.astype(self.dtype_s))

            def is_term(s):
                rel = s[self.relevant_indices]
                return any(self._box_contains(rel, lo, hi) for lo, hi in
                           zip(self.term_lows, self.term_highs))
            self.is_terminal_state = is_term

    def _box_contains(self, x, low, high):
        """gymnasium Box.contains."""
        if not isinstance(x, np.ndarray):
            x = np.asarray(x, dtype=low.dtype)
        return bool(np.can_cast(x.dtype, low.dtype) and x.shape == low.shape
                    and np.all(x >= low) and np.all(x <= high))

    def _box_sample(self, rng):
        """gymnasium Box.sample for the (homogeneous-bound) feature space."""
        if self.kind == "grid":
            hi = np.array(self.grid_shape, dtype=np.int64) + 1
            return np.floor(rng.uniform(low=np.zeros(len(hi)), high=hi,
                                        size=(len(hi),))).astype(np.int64)
        D = self.state_space_dim
        if np.isinf(self.state_space_max):
            sample = rng.normal(size=(D,))
        else:
            sample = rng.uniform(low=self.feat_low, high=self.feat_high,
                                 size=(D,))
        return sample.astype(self.dtype_s)

    def _init_init_state_dist(self):
        """rl_toy_env.py:992-1022."""
        if self.kind != "discrete":
            return
        cfg = self.config
        if not (self.use_custom_mdp and "init_state_dist" in cfg):
            A = self.action_space_size
            n_nonterm = A - self.num_terminal_states
            per_set = ([1 / (n_nonterm * self.diameter)] * n_nonterm
                       + [0] * self.num_terminal_states)
            cfg["relevant_init_state_dist"] = np.array(per_set * self.diameter)
        self.init_state_dist = np.asarray(cfg["relevant_init_state_dist"])
        self.init_cdf = normalised_cdf(self.init_state_dist)
        if self.irrelevant_features:  # :1024-1035, uniform over ALL states
            S1 = self.state_space_size_irr
            cfg["irrelevant_init_state_dist"] = np.array([1 / S1] * S1)
            self.init_cdf_irr = normalised_cdf(cfg["irrelevant_init_state_dist"])

    def _next_set_prob(self, s):
        """Uniform over the next independent set (:1074-1092)."""
        A, S = self.action_space_size, self.state_space_size
        i_s = s // A
        prob = np.zeros((S,))
        ind_1 = ((i_s + 1) * A) % S
        ind_2 = ((i_s + 2) * A) % S
        if ind_2 <= ind_1:
            ind_2 += S
        prob[ind_1:ind_2] = np.ones((A,)) / A
        return prob

    def _init_transition_function(self):
        """rl_toy_env.py:1042-1152 (S-stream draws)."""
        if self.kind != "discrete":
            return
        cfg = self.config
        if self.use_custom_mdp:
            P = cfg["transition_function"]
            assert not callable(P), "callable P: next row"
            self.transition_matrix = np.asarray(P)
        else:
            S, A = self.state_space_size, self.action_space_size
            P = np.zeros((S, A), dtype=object)
            P[:] = -1
            for s in range(S):
                if self.maximally_connected and self.diameter == 1:
                    P[s] = np.squeeze(self.rng_S.choice(
                        S, size=A, p=None, replace=False))
                elif self.maximally_connected:
                    P[s] = np.squeeze(self.rng_S.choice(
                        S, size=A, p=self._next_set_prob(s), replace=False))
                else:
                    prob = self._next_set_prob(s)
                    for a in range(A):
                        P[s, a] = int(np.squeeze(self.rng_S.choice(
                            S, size=1, p=prob, replace=True)))
            for i_s in range(self.diameter):  # terminal self-loops :1135-1148
                for s in range(A - self.num_terminal_states, A):
                    P[i_s * A + s, :] = i_s * A + s
            self.transition_matrix = P
            if self.irrelevant_features:
                self._init_transition_function_irr()
        # A1': the noisy-transition cdf depends only on (P[s,a], p)
        if self.kind == "discrete" and self.transition_noise:
            self.noise_cdf = self._noise_cdf(self.state_space_size)
            if self.irrelevant_features:
                self.noise_cdf_irr = self._noise_cdf(self.state_space_size_irr)

    def _noise_cdf(self, S):
        cdf = np.empty((S, S))
        for nxt in range(S):
            probs = np.ones((S,)) * self.transition_noise / (S - 1)
            probs[nxt] = 1 - self.transition_noise
            cdf[nxt] = normalised_cdf(probs)
        return cdf

    def _init_transition_function_irr(self):
        """rl_toy_env.py:1154-1230 (S' stream): like the relevant table, but
        the next-set probabilities are always passed (also for diameter 1) and
        there are no terminal self-loops."""
        S1, A1 = self.state_space_size_irr, self.action_space_size_irr
        P = np.zeros((S1, A1), dtype=object)
        P[:] = -1
        for s in range(S1):
            i_s = s // A1
            prob = np.zeros((S1,))
            ind_1 = ((i_s + 1) * A1) % S1
            ind_2 = ((i_s + 2) * A1) % S1
            if ind_2 <= ind_1:
                ind_2 += S1
            prob[ind_1:ind_2] = np.ones((A1,)) / A1
            if self.maximally_connected:
                P[s] = np.squeeze(self.rng_S1.choice(
                    S1, size=A1, p=prob, replace=False))
            else:
                for a in range(A1):
                    P[s, a] = int(np.squeeze(self.rng_S1.choice(
                        S1, size=1, p=prob, replace=True)))
        self.transition_matrix_irr = P

    def _sequences_with_repeats(self, n, length, fraction, diameter):
        """get_sequences(repeats=True) :1291-1338 (one E draw in total)."""
        A = self.action_space_size
        total = n ** length
        k = int(fraction * total) or 1
        nums = self.rng_E.choice(total, size=k, replace=False)
        out = []
        for i_s in range(diameter):
            for num in nums:
                seq = []
                while len(seq) != length:
                    seq.append(num % n + ((len(seq) + i_s) % diameter) * A)
                    num = num // n
                out.append(seq)
        return out

    def _sequences_without_repeats(self, n, length, fraction, diameter):
        """get_sequences(repeats=False) :1346-1452 (one E draw per set)."""
        A = self.action_space_size
        assert length <= diameter * n
        radices = [n - (i // diameter) for i in range(length)]
        out = []
        for i_s in range(diameter):
            total = np.prod(radices)
            k = int(fraction * total) or 1
            nums = self.rng_E.choice(total, size=k, replace=False)
            for num in nums:
                pools = [list(range(n)) for _ in range(diameter)]
                seq = []
                for pos, radix in enumerate(radices):
                    which = (pos + i_s) % diameter
                    rem = num % radix
                    seq.append(pools[which].pop(rem) + which * A)
                    num = num // radix
                assert seq not in out
                out.append(seq)
        return out

    def _init_reward_function(self):
        """rl_toy_env.py:1253-1567."""
        if self.kind != "discrete":
            return
        cfg = self.config
        if self.use_custom_mdp:
            R = cfg["reward_function"]
            assert not callable(R), "callable R: next row"
            self.reward_matrix = np.asarray(R)
            return
        n = self.action_space_size - self.num_terminal_states
        gen = (self._sequences_with_repeats if self.repeats_in_sequences
               else self._sequences_without_repeats)
        seqs = gen(n, self.sequence_length, self.reward_density, self.diameter)
        reward_dist = self.reward_dist
        assert not callable(reward_dist), "callable reward_dist: next row"
        rews = None
        if isinstance(reward_dist, list):  # :1528-1544
            num = self.diameter * len(seqs)
            rews = [1.0] if num == 1 else np.linspace(
                reward_dist[0], reward_dist[1], num=num)
            assert rews[-1] == 1.0
            self.rng_E.shuffle(rews)
        table = {}
        for seq in seqs:  # insert_sequence :1475-1504
            seq = tuple(seq)
            table[seq] = rews[len(table)] if rews is not None else 1.0
            if self.make_denser:  # q1: prefixes are inserted, never paid out
                for k in range(1, len(seq)):
                    sub = seq[:k]
                    if sub not in table:
                        table[sub] = 0.0
                    table[sub] += table[seq] * k / len(seq)
        self.rewardable_sequences = table

    # ---- reset ------------------------------------------------------------
    def reset(self, seed=None, options=None):
        """rl_toy_env.py:2217-2377."""
        if seed is not None:
            self.rng_E, _ = np_random(seed)
        self.reward_buffer = [0.0] * self.delay
        self.total_episodes += 1
        L1 = self.augmented_state_length - 1
        if self.kind == "discrete":
            u = self.draws.reset_uniform()
            s0 = choice_from_uniform(self.init_cdf, u)
            self.curr_state = s0
            if self.irrelevant_features:  # second E draw :2260-2264
                u1 = self.draws.irr_reset_uniform()
                self.curr_state = (s0, choice_from_uniform(self.init_cdf_irr, u1))
            self.augmented_state = [np.nan] * L1 + [s0]
        elif self.kind == "grid":
            # :2325-2345: Box(0, grid_shape).sample() of an int Box draws
            # floor(uniform(0, shape + 1)) -- the cell index `shape` itself
            # (one past the grid) can come out; no terminal rejection happens
            s0 = self.draws.reset_box(self)
            self.curr_state = np.asarray(s0).astype(self.dtype_s)
            self.augmented_state = [np.nan] * L1 + [
                list(self.curr_state[[0, 1]])]
        else:
            while True:
                s0 = self.draws.reset_box(self)
                if not self.is_terminal_state(s0):
                    break
            self.curr_state = s0
            zero = np.zeros((self.state_space_dim,), dtype=self.dtype_s)
            self.state_derivatives = [zero.copy()
                                      for _ in range(self.dynamics_order + 1)]
            self.state_derivatives[0] = s0.copy()
            self.augmented_state = [[np.nan] * self.state_space_dim
                                    for _ in range(L1)] + [s0.copy()]
        if self.image_representations:
            self.curr_obs = self._image(self.curr_state)
        else:
            self.curr_obs = self.curr_state
        self.curr_state = self.dtype_s(self.curr_state)
        self.curr_obs = self.dtype_o(self.curr_obs)
        self.reached_terminal = False
        self.total_abs_noise_in_reward_episode = 0
        self.total_abs_noise_in_transition_episode = (
            np.zeros((self.state_space_dim,)) if self.kind == "continuous"
            else None)
        self.total_noisy_transitions_episode = 0
        self.total_reward_episode = 0
        self.total_transitions_episode = 0
        return self.curr_obs, {}

    # ---- transition -------------------------------------------------------
    def _transition_discrete(self, state, action):
        """rl_toy_env.py:1602-1622."""
        nxt = int(self.transition_matrix[state, action])
        if self.transition_noise:  # falsy 0 / None => no draw (q2)
            u = self.draws.transition_uniform()
            new = choice_from_uniform(self.noise_cdf[nxt], u)
            if new != nxt:
                self.total_noisy_transitions_episode += 1
            nxt = new
        return nxt

    def _transition_continuous(self, state, action):
        """rl_toy_env.py:1625-1725."""
        assert action.shape == (self.state_space_dim,)
        sd = self.state_derivatives
        n = self.dynamics_order
        if self._box_contains(action, self.act_low, self.act_high):
            sd[-1] = action / self.inertia
            # scipy.special.factorial(arange(1, n+1)) is a float64 array, so
            # each term is (f32 * f32) / f64 -> f64, rounded back by the +=
            fact = np.array([math.factorial(k) for k in range(1, n + 1)],
                            dtype=np.float64)
            for i in range(n):
                for j in range(n - i):
                    sd[i] += sd[i + j + 1] * (self.time_unit ** (j + 1)) / fact[j]
            nxt = sd[0].copy()
        else:
            nxt = state  # frozen; derivatives kept (quirk list, hard part 5)
        if self.has_transition_noise and self.transition_noise is not None:
            noise = self.draws.state_normal(self.transition_noise,
                                            self.state_space_dim)
        else:
            noise = np.zeros(self.state_space_dim)
        self.total_abs_noise_in_transition_episode += np.abs(noise)
        nxt += noise  # q3: not written back into sd[0]
        self.noise_in_transition = noise
        if self.image_representations:
            in_bounds = False  # q4: ImageContinuous.contains(non-image) is None
        else:
            in_bounds = self._box_contains(nxt, self.feat_low, self.feat_high)
        if not in_bounds:
            nxt = np.clip(nxt, -self.state_space_max, self.state_space_max)
            zero = np.array([0.0] * self.state_space_dim, dtype=self.dtype_s)
            self.state_derivatives = [zero.copy() for _ in range(n + 1)]
            self.state_derivatives[0] = nxt.copy()
        if self.config["reward_function"] == "move_to_a_point":  # :1719-1725
            rel = np.array(nxt, dtype=self.dtype_s)[self.relevant_indices]
            if np.linalg.norm(rel - self.target_point) < self.target_radius:
                self.reached_terminal = True
        return nxt

    @staticmethod
    def _grid_action_valid(action):
        """GridActionSpace.contains (grid_action_space.py:25-39) and the dtype
        test of rl_toy_env.py:1730-1733."""
        x = np.array(action)
        if x.dtype.kind != "i" or x.dtype != np.int64:
            return False
        if not np.all((x == 0) | (x == 1) | (x == -1)):
            return False
        return int(np.sum(np.abs(x))) in (0, 1)

    def _transition_grid(self, state, action):
        """rl_toy_env.py:1727-1778."""
        nd = len(self.grid_shape)
        if self._grid_action_valid(action) and np.array(action).shape == (nd,):
            action = list(action)
            if self.transition_noise:
                if self.draws.grid_noise_uniform() < self.transition_noise:
                    while True:  # substitute a different action
                        ind, val = self.draws.grid_action_sample(nd)
                        new_action = [0] * nd
                        new_action[ind] = val
                        if new_action != [int(a) for a in action]:
                            self.total_noisy_transitions_episode += 1
                            action = new_action
                            break
            nxt = []
            for i in range(nd):
                v = int(state[i]) + int(action[i])
                v = max(v, 0)
                if v >= self.grid_shape[i]:
                    v = self.grid_shape[i] - 1
                nxt.append(v)
        else:
            nxt = [int(v) for v in state]  # noop (and a warning)
        rel = nxt[:nd // 2] if self.config.get("irrelevant_features") else nxt
        if list(self.target_point) == rel:
            self.reached_terminal = True
        return np.array(nxt)

    # ---- reward -----------------------------------------------------------
    def _reward(self, action):
        """rl_toy_env.py:1782-1990."""
        aug, d, L = self.augmented_state, self.delay, self.sequence_length
        reward = 0.0
        if self.use_custom_mdp:
            reward = self.reward_matrix[aug[-2], action]  # R(s, a) :1262-1264
        elif self.kind == "discrete":
            if not np.isnan(aug[d]):
                key = tuple(aug[1 + d:self.augmented_state_length])
                reward = self.rewardable_sequences.get(key, 0.0)
        elif self.kind == "grid":  # :1947-1965, Manhattan distances
            tgt = np.array(self.target_point)
            new = np.array(aug)[-1]
            if self.make_denser:
                old = np.array(aug)[-2]
                reward += np.abs(old - tgt).sum() - np.abs(new - tgt).sum()
            elif list(new) == list(self.target_point):
                reward += 1.0
        else:
            if not np.isnan(aug[d][0]) and \
                    self.config["reward_function"] == "move_along_a_line":
                # :1865-1910: fit a line (first right singular vector) through
                # the last sequence_length relevant states, reward = - mean
                # distance of those states from it.  The fp32 SVD gives the
                # direction; the end points (a float64 linspace times it) and
                # every distance are float64
                data = np.array(aug, dtype=self.dtype_s)[
                    1 + d:self.augmented_state_length, self.relevant_indices]
                mean = data.mean(axis=0)
                _, _, vv = np.linalg.svd(data - mean)
                ends = vv[0] * np.linspace(-1, 1, 2)[:, np.newaxis]
                ends += mean
                total = 0
                for pt in data:
                    total += dist_of_pt_from_line(pt, ends[0], ends[-1])
                reward += -total / self.sequence_length
            elif not np.isnan(aug[d][0]):
                window = np.array(aug, dtype=self.dtype_s)
                new_rel = window[-1, self.relevant_indices]
                if self.make_denser:
                    old_rel = window[-2, self.relevant_indices]
                    reward = -np.linalg.norm(new_rel - self.target_point)
                    reward += np.linalg.norm(old_rel - self.target_point)
                elif (np.linalg.norm(new_rel - self.target_point)
                      < self.target_radius):
                    reward = 1.0
                reward -= self.action_loss_weight * np.linalg.norm(
                    np.array(action, dtype=self.dtype_s))
        # shared tail :1968-1990
        self.reward_buffer.append(reward)
        reward = self.reward_buffer.pop(0)
        if self.total_transitions_episode % self.reward_every_n_steps != 0:
            reward = 0.0
        noise = (self.draws.reward_normal(self.reward_noise_std)
                 if self.has_reward_noise and self.reward_noise_std is not None
                 else 0)
        self.total_abs_noise_in_reward_episode += np.abs(noise)
        self.total_reward_episode += reward
        reward += noise
        reward *= self.reward_scale
        reward += self.reward_shift
        return reward

    # ---- step -------------------------------------------------------------
    def step(self, action):
        """rl_toy_env.py:1992-2125."""
        irr = self.kind == "discrete" and self.irrelevant_features
        if irr:  # :2029-2035
            state_irr, action_irr = self.curr_state[1], action[1]
            state, action = self.curr_state[0], action[0]
            nxt = self._transition_discrete(state, action)
        elif self.kind == "discrete":
            nxt = self._transition_discrete(self.curr_state, action)
        elif self.kind == "grid":
            nxt = self._transition_grid(self.curr_state, action)
        else:
            nxt = self._transition_continuous(self.curr_state, action)
        del self.augmented_state[0]
        if self.kind == "grid":  # :2055-2056, the relevant cell only
            self.augmented_state.append([nxt[i] for i in range(2)])
        else:
            self.augmented_state.append(nxt if self.kind == "discrete"
                                        else nxt.copy())
        self.total_transitions_episode += 1
        self.reward = self._reward(action)
        if irr:  # :2062-2082, after the reward (draw order S, E, S')
            nxt_irr = int(self.transition_matrix_irr[state_irr, action_irr])
            if self.transition_noise:
                u = self.draws.irr_transition_uniform()
                nxt_irr = choice_from_uniform(self.noise_cdf_irr[nxt_irr], u)
            nxt = (nxt, nxt_irr)
        obs = self._image(nxt) if self.image_representations else nxt
        self.curr_state = self.dtype_s(nxt)
        self.curr_obs = self.dtype_o(obs)
        self.done = bool(self.is_terminal_state(self.augmented_state[-1])
                         or self.reached_terminal)
        if self.done:
            self.reward += self.term_state_reward * self.reward_scale
        return self.curr_obs, self.reward, self.done, False, {
            "curr_state": self.curr_state}

    # ---- images -----------------------------------------------------------
    def _image(self, state):
        if self.kind == "discrete":
            if self.irrelevant_features:  # image_multi_discrete.py:272-288
                return np.atleast_3d(np.concatenate(
                    [self._image_discrete(int(s)) for s in state], axis=0))
            return np.atleast_3d(self._image_discrete(int(state)))
        if self.kind == "grid":
            return self._image_grid(np.asarray(state))
        return self._image_continuous(state)

    def image_params(self, state):
        """The random part of ImageMultiDiscrete.generate_image (:149-181,
        :251,:258-259) in draw order; returns dict R, shift_w, shift_h,
        rotation (None = no rotate), flip (0 none, 1 LR, 2 TB)."""
        W, H = self.image_width, self.image_height
        tr = self.image_transforms
        R = 20  # circle_radius passed at rl_toy_env.py:715
        shift_w, shift_h = int(W / 2), int(H / 2)
        if "scale" in tr:
            lo, hi = self.image_scale_range
            max_R = np.log(hi * R)
            min_R = np.log(lo * R)
            sample = np.exp(min_R + self.draws.image_scale_u() * (max_R - min_R))
            R = int(sample)
        if "shift" in tr:
            q = self.image_sh_quant
            mw, mh = W / 2 - R, H / 2 - R
            aw = self.draws.image_integers(-mw + 1, mw)
            ah = self.draws.image_integers(-mh + 1, mh)
            shift_w += (aw // q) * q
            shift_h += (ah // q) * q
        rotation = None
        if "rotate" in tr:
            rotation = self.draws.image_integers(360)
            rotation = (rotation // self.image_ro_quant) * self.image_ro_quant
        flip = 0
        if "flip" in tr:
            if self.draws.image_integers(2) == 0:
                flip = 1 if self.draws.image_integers(2) == 0 else 2
        return dict(R=R, shift_w=shift_w, shift_h=shift_h, rotation=rotation,
                    flip=flip)

    def _image_discrete(self, state):
        """image_multi_discrete.py:129-270 (default polygons, mode 'L')."""
        import PIL.Image as Image
        import PIL.ImageDraw as ImageDraw
        prm = self.image_params(state)
        self.last_image_params = prm
        sides = state + 3
        img = Image.new("L", (self.image_width, self.image_height))
        pts = []
        for i in range(sides):
            ang = (2 * np.pi / sides) * i
            pts.append((int(prm["shift_w"] + prm["R"] * np.cos(ang)),
                        int(prm["shift_h"] + prm["R"] * np.sin(ang))))
        ImageDraw.Draw(img).polygon(pts, fill=255)
        if prm["rotation"] is not None:
            img = img.rotate(prm["rotation"])
        if prm["flip"] == 1:
            img = img.transpose(Image.FLIP_LEFT_RIGHT)
        elif prm["flip"] == 2:
            img = img.transpose(Image.FLIP_TOP_BOTTOM)
        return np.array(img).T

    def _image_cont_setup(self):
        """ImageContinuous.__init__ (image_continuous.py:59-114) as built at
        rl_toy_env.py:767-776: relevant_indices defaults to [0, 1]."""
        D = self.state_space_dim
        self.img_rel = [0, 1]
        self.img_irr = sorted(set(range(D)) - set(self.img_rel))
        assert len(self.img_irr) <= 2
        self.target_pixel = self._to_pixel(self.target_point)

    def _to_pixel(self, vec):
        """image_continuous.py:248-277: f32 ratio, promoted to f64 by the
        int64 shape tuple, truncated."""
        hi = self.feat_high[self.img_rel]
        lo = self.feat_low[self.img_rel]
        frac = (vec - lo) / (hi - lo)
        return (frac * (self.image_width, self.image_height)).astype(int)

    def _image_continuous_one(self, pos, relevant):
        """image_continuous.py:116-208 (no grid)."""
        import PIL.Image as Image
        import PIL.ImageDraw as ImageDraw
        img = Image.new("RGB", (self.image_width, self.image_height),
                        color=(208, 208, 208))
        draw = ImageDraw.Draw(img)
        if relevant:
            for lo, hi in zip(self.term_lows, self.term_highs):
                draw.rectangle([tuple(self._to_pixel(lo)),
                                tuple(self._to_pixel(hi))], fill=(0, 0, 0))
            tp = self.target_pixel
            draw.ellipse([tuple(tp - 5), tuple(tp + 5)], fill=(0, 255, 0))
        pp = self._to_pixel(pos)
        draw.ellipse([tuple(pp - 5), tuple(pp + 5)], fill=(0, 0, 255))
        return np.transpose(np.array(img), axes=(1, 0, 2))

    def _grid_pixel(self, vec):
        """convert_to_pixel (image_continuous.py:248-277) over the int64 Box
        [0, grid_shape] of the first two dimensions (also for the irrelevant
        sub-image: "both sub-spaces have the same max and min")."""
        hi = np.array(self.grid_shape[:2], dtype=np.int64)
        frac = (np.asarray(vec) - 0) / (hi - 0)
        return (frac * (self.image_width, self.image_height)).astype(int)

    def _image_grid_one(self, pos, relevant):
        """image_continuous.py:116-208 with draw_grid (grid lines, terminal
        cells, target and agent discs at the cell centres)."""
        import PIL.Image as Image
        import PIL.ImageDraw as ImageDraw
        W, H = self.image_width, self.image_height
        img = Image.new("RGB", (W, H), color=(208, 208, 208))
        draw = ImageDraw.Draw(img)
        off = 0 if relevant else 2
        gs = self.grid_shape
        for i in range(1, gs[0 + off] + 1):
            x = i * W // gs[0 + off] - 1
            draw.line([(x, H), (x, 0)], fill=(255, 255, 255))
        for j in range(1, gs[1 + off]):
            y = j * H // gs[0 + off]  # (sic: the first extent, :152)
            draw.line([(W, y), (0, y)], fill=(255, 255, 255))
        if relevant:
            for cell in self.term_cells:
                c = np.array(cell, dtype=np.float64).astype(np.int64)
                draw.rectangle([tuple(self._grid_pixel(c)),
                                tuple(self._grid_pixel(c + 1.0))], fill=(0, 0, 0))
            tp = self._grid_pixel(np.array(self.target_point, dtype=float) + 0.5)
            draw.ellipse([tuple(tp - 5), tuple(tp + 5)], fill=(0, 255, 0))
        pp = self._grid_pixel(np.asarray(pos).astype(float) + 0.5)
        draw.ellipse([tuple(pp - 5), tuple(pp + 5)], fill=(0, 0, 255))
        return np.transpose(np.array(img), axes=(1, 0, 2))

    def _image_grid(self, obs):
        parts = [self._image_grid_one(obs[[0, 1]], True)]
        if len(self.grid_shape) == 4:
            parts.append(self._image_grid_one(obs[[2, 3]], False))
        return np.atleast_3d(np.concatenate(parts, axis=0))

    def _image_continuous(self, obs):
        """image_continuous.py:210-246."""
        parts = [self._image_continuous_one(obs[self.img_rel], True)]
        if self.img_irr:
            parts.append(self._image_continuous_one(obs[self.img_irr], False))
        return np.atleast_3d(np.concatenate(parts, axis=0))
