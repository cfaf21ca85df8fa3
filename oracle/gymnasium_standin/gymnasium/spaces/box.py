"""gymnasium.spaces.Box restated (sampling + contains semantics of 1.x)."""
import numpy as np

from .space import Space


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        assert dtype is not None
        self.dtype = np.dtype(dtype)
        if shape is not None:
            shape = tuple(int(d) for d in shape)
        elif isinstance(low, np.ndarray):
            shape = low.shape
        elif isinstance(high, np.ndarray):
            shape = high.shape
        elif np.isscalar(low) and np.isscalar(high):
            shape = (1,)
        else:
            raise ValueError("Box shape is not specified")
        with np.errstate(invalid="ignore", over="ignore"):
            low_arr = np.full(shape, low, dtype=float) if np.isscalar(low) \
                else np.asarray(low, dtype=float)
            high_arr = np.full(shape, high, dtype=float) if np.isscalar(high) \
                else np.asarray(high, dtype=float)
        self.bounded_below = -np.inf < low_arr
        self.bounded_above = np.inf > high_arr
        if self.dtype.kind in "iu":
            info = np.iinfo(self.dtype)
            low_arr = np.where(np.isneginf(low_arr), info.min, low_arr)
            high_arr = np.where(np.isposinf(high_arr), info.max, high_arr)
        self._shape = shape
        self.low = low_arr.astype(self.dtype)
        self.high = high_arr.astype(self.dtype)
        super().__init__(self._shape, self.dtype, seed)

    def is_bounded(self, manner="both"):
        below = bool(np.all(self.bounded_below))
        above = bool(np.all(self.bounded_above))
        return {"both": below and above, "below": below, "above": above}[manner]

    def sample(self, mask=None):
        high = self.high if self.dtype.kind == "f" \
            else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp_bounded = ~self.bounded_below & self.bounded_above
        low_bounded = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(
            size=unbounded[unbounded].shape)
        sample[low_bounded] = (self.np_random.exponential(
            size=low_bounded[low_bounded].shape) + self.low[low_bounded])
        sample[upp_bounded] = (-self.np_random.exponential(
            size=upp_bounded[upp_bounded].shape) + self.high[upp_bounded])
        sample[bounded] = self.np_random.uniform(
            low=self.low[bounded], high=high[bounded],
            size=bounded[bounded].shape)
        if self.dtype.kind in ["i", "u", "b"]:
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(np.can_cast(x.dtype, self.dtype)
                    and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
