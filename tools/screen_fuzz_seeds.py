"""Which seeds of tests/fuzz_configs.py does the UNMODIFIED reference accept?
Build container only (needs /root/reference).  Prints the seed lists to paste
into fuzz_configs.py; a configuration the reference rejects (or that makes it
raise within 40 steps) is no test case."""
import copy
import signal
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
from oracle.ref_loader import make_reference_env  # noqa: E402
from tests import fuzz_configs as fz  # noqa: E402
from tests.golden.cases import materialise  # noqa: E402


class _Timeout(Exception):
    pass


def _alarm(*_):
    raise _Timeout()


def accepted(cfg):
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(20)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = make_reference_env(materialise(copy.deepcopy(cfg)))
            env.reset(seed=1)
            for _ in range(40):
                a = env.action_space.sample()
                if cfg["state_space_type"] == "grid":
                    a = [int(x) for x in a]  # the reference compares lists
                _, _, done, _, _ = env.step(a)
                if done:
                    env.reset()
        return True, ""
    except _Timeout:
        return False, "timeout"
    except Exception as e:  # noqa: BLE001
        return False, "%s: %s" % (type(e).__name__, str(e)[:80])
    finally:
        signal.alarm(0)


for name, gen, n in (("DISCRETE", fz.discrete_fuzz_config, 60),
                     ("CONTINUOUS", fz.continuous_fuzz_config, 60),
                     ("GRID", fz.grid_fuzz_config, 40),
                     ("IMAGE", fz.image_fuzz_config, 40)):
    ok = []
    for s in range(n):
        good, why = accepted(gen(s))
        if good:
            ok.append(s)
        else:
            print("# %s seed %d rejected by the reference: %s" % (name, s, why))
    print("%s_SEEDS = %r" % (name, ok))
