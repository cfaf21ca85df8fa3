from .space import Space  # noqa: F401
from .discrete import Discrete  # noqa: F401
from .box import Box  # noqa: F401
from .multi_discrete import MultiDiscrete  # noqa: F401
from .tuple import Tuple  # noqa: F401
