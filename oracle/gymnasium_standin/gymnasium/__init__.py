"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

gymnasium is not installed in the build image and cannot be fetched (no
network).  The reference (`/root/reference/mdp_playground`) imports it at
module top, so this stand-in provides exactly the classes the hot-path files
touch, restating gymnasium 1.x's published behaviour:

  * `utils.seeding.np_random`  -> Generator(PCG64(SeedSequence(seed)))
  * `Env.reset(seed=)`         -> re-seeds `_np_random` iff a seed is given
  * `spaces.Space/Discrete/Box/MultiDiscrete/Tuple`
  * `envs.registration.register`, `error.DependencyNotInstalled`

It is only ever put on `sys.path` by `oracle/ref_loader.py` (golden-vector
generation and oracle-vs-reference validation in the build container).  The
product package never imports it.
"""
from . import error, utils, spaces  # noqa: F401
from .core import Env, Wrapper  # noqa: F401
from .envs.registration import register, make  # noqa: F401

__version__ = "1.0.0-standin"
