"""256-layer ziggurat for N(0,1) (TEST INFRASTRUCTURE ONLY).

CPU restatement of the algorithm behind `numpy.random.Generator.normal`,
which is what the reference's reward noise ends up calling
(`rl_toy_env.py:403`, `self.np_random.normal(0, sigma)` at `:1982`): numpy
1.17+ `random_standard_normal` (numpy/random/src/distributions/
distributions.c, pinned by the image's numpy 2.3): one 64-bit word r per
attempt,

    idx = r & 0xff; sign = (r >> 8) & 1; rabs = (r >> 9) & (2^52 - 1)
    x = rabs * wi[idx] (negated if sign); accept if rabs < ki[idx]   (99.3 %)
    idx == 0: tail beyond R by Marsaglia's method, two doubles per attempt
    else    : accept if (fi[idx-1] - fi[idx]) * U + fi[idx] < exp(-x^2 / 2)
    otherwise start over with a fresh word

with doubles U = (word >> 11) * 2^-53.  The tables are numpy's own
(`oracle/ziggurat_tables.npz`, written by tools/extract_numpy_ziggurat.py;
Marsaglia & Tsang 2000 construction, R = 3.65415288536..., 256 layers, 52-bit
scale).  `tests/test_ziggurat.py` pins both the tables and this restatement:
PCG64's raw words fed through `standard_normal_stream` reproduce
`Generator.standard_normal` BIT FOR BIT, and `make_tables_highprec` (the
construction redone with 60-digit arithmetic) agrees with the shipped tables
to the few ulps of error numpy's tables carry.  The CUDA kernels
(csrc/ziggurat.cuh) use the same tables (compiled into the library from
csrc/ziggurat_tables.h, same generator script; compared in the CPU tests) fed
with Philox words -- the product never imports oracle/.
"""
import math
import os

import numpy as np

N_LAYERS = 256
ZIG_R = 3.6541528853610087963519472518
ZIG_INV_R = 0.27366123732975827203338247596
_M52 = (1 << 52) - 1


def make_tables_highprec():
    """The table construction with 60-digit arithmetic: (ki, wi, fi).  Only a
    cross-check of the shipped tables (which carry a few ulps of error)."""
    import mpmath as mp
    mp.mp.dps = 60
    r = mp.mpf("3.6541528853610087963519472518")
    f = lambda x: mp.exp(-x * x / 2)
    # area of every layer = base strip (rectangle r * f(r) + tail)
    v = r * f(r) + mp.sqrt(mp.pi / 2) * mp.erfc(r / mp.sqrt(2))
    m = mp.mpf(2) ** 52
    # layers are indexed like Marsaglia-Tsang's zigset: idx 0 = base strip,
    # idx 255 = the widest proper layer (right edge r), idx 1 the topmost
    # idx i (>= 1) has right edge e[i]; e[255] = r, decreasing towards idx 1
    e = [None] * N_LAYERS
    e[N_LAYERS - 1] = r
    for i in range(N_LAYERS - 2, 0, -1):
        e[i] = mp.sqrt(-2 * mp.log(v / e[i + 1] + f(e[i + 1])))
    ki = np.zeros(N_LAYERS, dtype=np.uint64)
    wi = np.zeros(N_LAYERS, dtype=np.float64)
    fi = np.zeros(N_LAYERS, dtype=np.float64)
    q = v / f(r)
    ki[0] = int(mp.floor(r / q * m))
    ki[1] = 0
    wi[0] = float(q / m)
    wi[N_LAYERS - 1] = float(r / m)
    fi[0] = 1.0
    fi[N_LAYERS - 1] = float(f(r))
    for i in range(N_LAYERS - 2, 0, -1):
        ki[i + 1] = int(mp.floor(e[i] / e[i + 1] * m))
        fi[i] = float(f(e[i]))
        wi[i] = float(e[i] / m)
    return ki, wi, fi


_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                 "ziggurat_tables.npz"))
        _TABLES = (z["ki"], z["wi"], z["fi"])
    return _TABLES


def _double(word):
    return (int(word) >> 11) * (1.0 / 9007199254740992.0)


def standard_normal_stream(next_word):
    """One N(0,1) draw consuming 64-bit words from `next_word()` exactly in
    numpy's order (scalar; the reference semantics)."""
    ki, wi, fi = tables()
    while True:
        r = int(next_word())
        idx = r & 0xFF
        r >>= 8
        sign = r & 1
        rabs = (r >> 1) & _M52
        x = rabs * float(wi[idx])
        if sign:
            x = -x
        if rabs < int(ki[idx]):
            return x
        if idx == 0:
            while True:
                xx = -ZIG_INV_R * math.log1p(-_double(next_word()))
                yy = -math.log1p(-_double(next_word()))
                if yy + yy > xx * xx:
                    return -(ZIG_R + xx) if (rabs >> 8) & 1 else ZIG_R + xx
        else:
            if ((float(fi[idx - 1]) - float(fi[idx])) * _double(next_word())
                    + float(fi[idx])) < math.exp(-0.5 * x * x):
                return x


def fast_path(words):
    """Vectorised first attempt: (x, accepted) for uint64 `words`."""
    ki, wi, _ = tables()
    w = np.asarray(words, dtype=np.uint64)
    idx = (w & np.uint64(0xFF)).astype(np.int64)
    sign = ((w >> np.uint64(8)) & np.uint64(1)).astype(bool)
    rabs = (w >> np.uint64(9)) & np.uint64(_M52)
    x = rabs.astype(np.float64) * wi[idx]
    x = np.where(sign, -x, x)
    return x, rabs < ki[idx]


def standard_normal_counter(main_words, retry_words):
    """Counter-based use (the kernels' scheme): `main_words` uint64[...] is the
    first attempt of every draw; a draw whose first attempt is rejected
    continues numpy's algorithm on its own word sequence
    `retry_words(flat_index, k)` -> uint64, k = 0, 1, 2, ..."""
    main = np.asarray(main_words, dtype=np.uint64)
    x, ok = fast_path(main)
    x = x.copy()
    flat_x, flat_main = x.reshape(-1), main.reshape(-1)
    for i in np.nonzero(~ok.reshape(-1))[0]:
        state = {"k": -1}

        def nxt(i=i, state=state):
            state["k"] += 1
            return flat_main[i] if state["k"] == 0 else retry_words(i, state["k"] - 1)
        flat_x[i] = standard_normal_stream(nxt)
    return x
