"""gymnasium.Env restated: only seeding behaviour matters to the reference."""
from .utils import seeding


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    action_space = None
    observation_space = None
    _np_random = None
    _np_random_seed = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
