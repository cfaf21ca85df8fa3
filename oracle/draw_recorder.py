"""Delegating proxy around a numpy Generator that logs every draw.

Installed on a *reference* env (`env._np_random`, `space._np_random`) to
capture the draws that the replay mode of the CUDA path and of the oracle
consume.  For `choice(p=...)` the underlying uniform is
recovered by replaying the bit generator: numpy's `Generator.choice` with
probabilities draws exactly one `random()` and returns
`searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')`; the proxy asserts
that identity on every call.
"""
import copy

import numpy as np


class RecordingGenerator:
    def __init__(self, rng, log, tag):
        self._rng = rng
        self._log = log
        self._tag = tag

    # -- draws we care about -------------------------------------------------
    def choice(self, a, size=None, replace=True, p=None, **kw):
        if p is not None and replace and (size is None or size == 1):
            shadow = np.random.Generator(copy.deepcopy(self._rng.bit_generator))
            out = self._rng.choice(a, size=size, replace=replace, p=p, **kw)
            u = shadow.random()
            cdf = np.cumsum(np.asarray(p, dtype=np.float64))
            cdf /= cdf[-1]
            idx = int(np.searchsorted(cdf, u, side="right"))
            assert idx == int(np.squeeze(out)), "choice(p) identity violated"
            assert shadow.bit_generator.state["state"] == \
                self._rng.bit_generator.state["state"], "choice(p) drew != 1"
            self._log.append((self._tag, "choice_u", u, idx))
            return out
        out = self._rng.choice(a, size=size, replace=replace, p=p, **kw)
        self._log.append((self._tag, "choice_other", None, np.array(out)))
        return out

    def normal(self, loc=0.0, scale=1.0, size=None):
        out = self._rng.normal(loc, scale, size)
        self._log.append((self._tag, "normal", (loc, scale), np.array(out)))
        return out

    def random(self, size=None, **kw):
        out = self._rng.random(size, **kw)
        self._log.append((self._tag, "random", None, np.array(out)))
        return out

    def integers(self, low, high=None, size=None, **kw):
        out = self._rng.integers(low, high, size, **kw)
        self._log.append((self._tag, "integers", (low, high), np.array(out)))
        return out

    def uniform(self, low=0.0, high=1.0, size=None):
        out = self._rng.uniform(low, high, size)
        self._log.append((self._tag, "uniform", None, np.array(out)))
        return out

    def __getattr__(self, name):
        return getattr(self._rng, name)


def install(env):
    """Wrap every generator the reference env draws from; return the log."""
    log = []
    env._np_random = RecordingGenerator(env._np_random, log, "env")
    for i, sp in enumerate(getattr(env, "observation_spaces", [])):
        sp.np_random  # force lazy init
        sp._np_random = RecordingGenerator(sp._np_random, log, f"obs{i}")
    if getattr(env, "image_representations", False):
        sp = env.observation_space
        sp.np_random
        if not isinstance(sp._np_random, RecordingGenerator):
            sp._np_random = RecordingGenerator(sp._np_random, log, "image")
    if env.config.get("state_space_type") == "grid":  # GridActionSpace.sample
        sp = env.action_space
        sp.np_random
        if not isinstance(sp._np_random, RecordingGenerator):
            sp._np_random = RecordingGenerator(sp._np_random, log, "action")
    fs = getattr(env, "feature_space", None)
    if fs is not None:
        fs.np_random
        if not isinstance(fs._np_random, RecordingGenerator):
            fs._np_random = RecordingGenerator(fs._np_random, log, "feature")
    return log
