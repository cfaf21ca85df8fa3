"""Host-side generation of the discrete MDP tables with the reference's seeded
numpy procedure, so that P, the rewardable sequences, the terminal set and
the initial-state distribution are identical to the reference's for the same
config (north_star: "Table and sequence generation stays on host").

Draw order, all on numpy Generator(PCG64):
  S stream (seed_dict["relevant_state_space"]): one choice() per P row
      (init_transition_function, rl_toy_env.py:1065-1132)
  E stream (seed_dict["env"], after the 7 fan-out draws): one choice() per
      independent set for the sequences (:1398-1402 / :1301-1305), then an
      optional shuffle of the reward values (:1539)
"""
from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np

from .config import EnvSpec, np_random


@dataclass
class DiscreteTables:
    n_states: int
    n_actions: int
    transition: np.ndarray        # int32 [S, A]
    terminal_states: np.ndarray   # int64 [n_term] (reference order)
    terminal_mask: np.ndarray     # uint8 [S]
    init_state_dist: np.ndarray   # float64 [S]
    init_cdf: np.ndarray          # float64 [S]
    noise_cdf: Optional[np.ndarray]   # float64 [S, S] or None
    sequences: np.ndarray         # int32 [n_seq, L] (full-length only)
    sequence_rewards: np.ndarray  # float64 [n_seq]
    rewardable_sequences: Dict[tuple, float]  # incl. dead make_denser prefixes
    reward_matrix: Optional[np.ndarray]  # float64 [S, A] (custom MDP)
    # irrelevant sub-MDP (irrelevant_features): own P, uniform start, own noise
    n_states_irr: int = 0
    n_actions_irr: int = 0
    transition_irr: Optional[np.ndarray] = None      # int32 [S1, A1]
    init_cdf_irr: Optional[np.ndarray] = None        # float64 [S1]
    noise_cdf_irr: Optional[np.ndarray] = None       # float64 [S1, S1]


def normalised_cdf(p):
    """The cdf numpy's Generator.choice(p=) searches: cumsum, / last."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


def _terminal_states(sp: EnvSpec):
    """init_terminal_states, discrete branch (:857-889)."""
    cfg = sp.config
    if sp.use_custom_mdp and "terminal_states" in cfg:
        ts = cfg["terminal_states"]
        if callable(ts):
            # a callable over a finite state set is just a mask
            ts = [s for s in range(sp.state_space_size) if ts(s)]
        return np.array(ts, dtype=np.int64).reshape(-1), None
    A = sp.action_space_size
    n_term = int(sp.terminal_state_density * A)
    states = np.array([j * A - 1 - i for j in range(1, sp.diameter + 1)
                       for i in range(n_term)], dtype=np.int64)
    cfg["terminal_states"] = states
    return states, n_term


def _init_state_dist(sp: EnvSpec, n_term):
    """init_init_state_dist (:995-1018)."""
    cfg = sp.config
    if sp.use_custom_mdp and "init_state_dist" in cfg:
        return np.asarray(cfg["relevant_init_state_dist"], dtype=np.float64)
    A = sp.action_space_size
    n_nonterm = A - n_term
    per_set = [1 / (n_nonterm * sp.diameter)] * n_nonterm + [0] * n_term
    dist = np.array(per_set * sp.diameter)
    cfg["relevant_init_state_dist"] = dist
    return dist


def _next_set_probabilities(s, S, A):
    """Uniform over the independent set that follows the one `s` is in."""
    i_s = s // A
    prob = np.zeros((S,))
    first = ((i_s + 1) * A) % S
    last = ((i_s + 2) * A) % S
    if last <= first:
        last += S
    prob[first:last] = np.ones((A,)) / A
    return prob


def sub_space_rngs(sp: EnvSpec):
    """Generators of observation_spaces[0] / [1] at table-generation time.
    Without images the two sub-spaces of an irrelevant_features env are wrapped
    in a TupleExtended seeded with an int (:738-741), and gymnasium 1.x
    Tuple.seed re-seeds every sub-space with integers(int32.max, size=2) of
    its own stream (parity of this cascade is unpinned by the reference,
    SURVEY.md 8c; it matches oracle/gymnasium_standin)."""
    sd = sp.seed_dict
    if not sp.irrelevant_features:
        return np_random(sd["relevant_state_space"])[0], None
    if sp.image_representations:
        return (np_random(sd["relevant_state_space"])[0],
                np_random(sd["irrelevant_state_space"])[0])
    rng_t, _ = np_random(sd["state_space"])
    sub = rng_t.integers(np.iinfo(np.int32).max, size=2)
    return np_random(int(sub[0]))[0], np_random(int(sub[1]))[0]


def _transition_matrix_irr(sp: EnvSpec):
    """init_transition_function, irrelevant part (:1154-1230): the next-set
    probabilities are always passed to choice() and no state self-loops."""
    S1, A1 = sp.state_space_size_irr, sp.action_space_size_irr
    rng = sub_space_rngs(sp)[1]
    P = np.full((S1, A1), -1, dtype=np.int64)
    for s in range(S1):
        p = _next_set_probabilities(s, S1, A1)
        if sp.maximally_connected:
            P[s] = np.squeeze(rng.choice(S1, size=A1, p=p, replace=False))
        else:
            for a in range(A1):
                P[s, a] = int(np.squeeze(rng.choice(S1, size=1, p=p)))
    return P.astype(np.int32)


def _noise_cdf(p, S):
    """Row s' = cdf of the noisy redraw around s' (:1605-1612)."""
    cdf = np.empty((S, S), dtype=np.float64)
    for nxt in range(S):
        probs = np.ones((S,)) * p / (S - 1)
        probs[nxt] = 1 - p
        cdf[nxt] = normalised_cdf(probs)
    return cdf


def _transition_matrix(sp: EnvSpec, n_term):
    """init_transition_function (:1046-1152)."""
    cfg = sp.config
    S, A = sp.state_space_size, sp.action_space_size
    if sp.use_custom_mdp:
        P = cfg["transition_function"]
        if callable(P):
            P = [[P(s, a) for a in range(A)] for s in range(S)]
        P = np.asarray(P, dtype=np.int64)
        assert P.shape == (S, A), "custom transition_function must be [S, A]"
        return P.astype(np.int32)
    rng = sub_space_rngs(sp)[0]
    P = np.full((S, A), -1, dtype=np.int64)
    for s in range(S):
        if sp.maximally_connected:
            p = None if sp.diameter == 1 else _next_set_probabilities(s, S, A)
            P[s] = np.squeeze(rng.choice(S, size=A, p=p, replace=False))
        else:
            p = _next_set_probabilities(s, S, A)
            for a in range(A):
                P[s, a] = int(np.squeeze(rng.choice(S, size=1, p=p)))
    for i_s in range(sp.diameter):  # terminal states self-loop (:1135-1148)
        for s in range(A - n_term, A):
            P[i_s * A + s, :] = i_s * A + s
    return P.astype(np.int32)


def _select_sequences(sp: EnvSpec, rng, n_nonterm):
    """get_sequences (:1273-1473): decode sampled sequence numbers."""
    A, L, diam = sp.action_space_size, sp.sequence_length, sp.diameter
    frac = sp.reward_density
    out = []
    if sp.repeats_in_sequences:  # mixed radix, digits may repeat
        total = n_nonterm ** L
        picked = rng.choice(total, size=int(frac * total) or 1, replace=False)
        for i_s in range(diam):
            for num in picked:
                seq = []
                for pos in range(L):
                    seq.append(int(num % n_nonterm) + ((pos + i_s) % diam) * A)
                    num = num // n_nonterm
                out.append(seq)
        return out
    assert L <= diam * n_nonterm, (
        "When there are no repeats in sequences, the sequence length should "
        "be <= diameter * maximum.")
    radices = [n_nonterm - (i // diam) for i in range(L)]
    for i_s in range(diam):  # factorial-number-system decode, no repeats
        total = np.prod(radices)
        picked = rng.choice(total, size=int(frac * total) or 1, replace=False)
        for num in picked:
            remaining = [list(range(n_nonterm)) for _ in range(diam)]
            seq = []
            for pos, radix in enumerate(radices):
                which = (pos + i_s) % diam
                seq.append(remaining[which].pop(int(num % radix)) + which * A)
                num = num // radix
            assert seq not in out
            out.append(seq)
    return out


def build_discrete_tables(sp: EnvSpec) -> DiscreteTables:
    cfg = sp.config
    S, A, L = sp.state_space_size, sp.action_space_size, sp.sequence_length
    term_states, n_term = _terminal_states(sp)
    mask = np.zeros((S,), dtype=np.uint8)
    mask[term_states] = 1
    init_dist = _init_state_dist(sp, n_term)
    assert init_dist.shape == (S,)
    P = _transition_matrix(sp, n_term)

    table: Dict[tuple, float] = {}
    sequences = np.zeros((0, L), dtype=np.int32)
    seq_rewards = np.zeros((0,), dtype=np.float64)
    reward_matrix = None
    if sp.use_custom_mdp:
        R = cfg["reward_function"]
        if callable(R):
            raise NotImplementedError(
                "callable custom reward_function cannot run on the device; "
                "pass an [S, A] array")
        reward_matrix = np.ascontiguousarray(np.asarray(R, dtype=np.float64))
        assert reward_matrix.shape == (S, A)
    else:
        rng = sp.env_rng  # E stream, positioned after the seed fan-out
        seqs = _select_sequences(sp, rng, A - n_term)
        values = None
        if isinstance(sp.reward_dist, list):  # :1528-1544
            num = sp.diameter * len(seqs)
            values = [1.0] if num == 1 else np.linspace(
                sp.reward_dist[0], sp.reward_dist[1], num=num)
            assert values[-1] == 1.0
            rng.shuffle(values)
        for seq in seqs:  # insert_sequence :1475-1504
            seq = tuple(seq)
            table[seq] = values[len(table)] if values is not None else 1.0
            if sp.make_denser:
                # Prefix entries are created like the reference does, but the
                # step path can never pay them out: the lookup key always has
                # length L (:1837-1841).  Kept for `rewardable_sequences`
                # equality only (SURVEY.md finding 3).
                for k in range(1, len(seq)):
                    sub = seq[:k]
                    table.setdefault(sub, 0.0)
                    table[sub] += table[seq] * k / len(seq)
        full = [(k, v) for k, v in table.items() if len(k) == L]
        sequences = np.array([k for k, _ in full], dtype=np.int32).reshape(-1, L)
        seq_rewards = np.array([float(v) for _, v in full], dtype=np.float64)

    noise_cdf = None
    if sp.transition_noise:  # falsy => the reference draws nothing (:1604)
        noise_cdf = _noise_cdf(sp.transition_noise, S)
    irr = {}
    if sp.irrelevant_features:
        S1 = sp.state_space_size_irr
        cfg["irrelevant_init_state_dist"] = np.array([1 / S1] * S1)  # :1024-1035
        irr = dict(
            n_states_irr=S1, n_actions_irr=sp.action_space_size_irr,
            transition_irr=np.ascontiguousarray(_transition_matrix_irr(sp)),
            init_cdf_irr=normalised_cdf(cfg["irrelevant_init_state_dist"]),
            noise_cdf_irr=(_noise_cdf(sp.transition_noise, S1)
                           if sp.transition_noise else None))

    return DiscreteTables(**irr, 
        n_states=S, n_actions=A, transition=np.ascontiguousarray(P),
        terminal_states=term_states, terminal_mask=mask,
        init_state_dist=init_dist, init_cdf=normalised_cdf(init_dist),
        noise_cdf=noise_cdf, sequences=np.ascontiguousarray(sequences),
        sequence_rewards=seq_rewards, rewardable_sequences=table,
        reward_matrix=reward_matrix)
