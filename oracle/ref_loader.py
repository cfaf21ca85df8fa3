"""Import the UNMODIFIED reference `mdp_playground` from /root/reference.

Looks in /root/reference (the build container) and then in oracle/_ref/ (the
git-ignored copy made by oracle/make_ref.py, which travels to the GPU box).
Test infrastructure: only tests/, smoke() and bench.py's CPU legs use it.
gymnasium is absent from the image, so `oracle/gymnasium_standin` is put on
sys.path first (see its docstring for what it restates).
"""
import contextlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MDPP_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "mdp_playground")):
    # the GPU box: the git-ignored copy made by oracle/make_ref.py travels there
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")
_STANDIN = os.path.join(_HERE, "gymnasium_standin")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mdp_playground"))


def load_reference():
    """Return the reference's RLToyEnv class (imports are cached)."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at " + REFERENCE_ROOT)
    try:
        import gymnasium  # noqa: F401  (a real install wins if present)
    except ModuleNotFoundError:
        if _STANDIN not in sys.path:
            sys.path.insert(0, _STANDIN)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from mdp_playground.envs.rl_toy_env import RLToyEnv
    return RLToyEnv


def make_reference_env(config):
    """Construct a reference env with its constructor chatter silenced."""
    import copy
    import logging
    import warnings
    cls = load_reference()
    logging.disable(logging.WARNING)  # the ctor logs at WARNING (:336,:883)
    try:
        with contextlib.redirect_stdout(io.StringIO()), \
                warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = cls(**copy.deepcopy(config))
    finally:
        logging.disable(logging.NOTSET)
    env.logger.setLevel(logging.ERROR)
    return env


def load_reference_wrapper():
    """The reference's GymEnvWrapper class (envs/gym_env_wrapper.py).  Its
    module imports Atari machinery at the top (`gymnasium.wrappers.
    AtariPreprocessing`, `ale_py`, `gym.register_envs`) that the wrapper tail
    never touches: stubbed here, for this import only."""
    import importlib
    import types
    load_reference()
    import gymnasium
    saved = {k: sys.modules.get(k) for k in ("gymnasium.wrappers", "ale_py")}
    wrappers = types.ModuleType("gymnasium.wrappers")
    wrappers.AtariPreprocessing = type("AtariPreprocessing", (), {})
    sys.modules["gymnasium.wrappers"] = wrappers
    sys.modules["ale_py"] = types.ModuleType("ale_py")
    had = hasattr(gymnasium, "register_envs")
    if not had:
        gymnasium.register_envs = lambda module: None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module("mdp_playground.envs.gym_env_wrapper")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if not had:
            del gymnasium.register_envs
    return mod.GymEnvWrapper
