import numpy as np

from .space import Space


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def seed(self, seed=None):
        if seed is None:
            return tuple(space.seed(None) for space in self.spaces)
        if isinstance(seed, int):
            super().seed(seed)
            subseeds = self.np_random.integers(
                np.iinfo(np.int32).max, size=len(self.spaces))
            return tuple(space.seed(int(s))
                         for space, s in zip(self.spaces, subseeds))
        if isinstance(seed, (tuple, list)):
            return tuple(space.seed(s) for space, s in zip(self.spaces, seed))
        raise TypeError(f"unsupported seed type {type(seed)}")

    def sample(self, mask=None):
        return tuple(space.sample() for space in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return (isinstance(x, tuple) and len(x) == len(self.spaces)
                and all(s.contains(p) for s, p in zip(self.spaces, x)))

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)
