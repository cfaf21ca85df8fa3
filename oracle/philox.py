"""numpy restatement of mdp_playground_b200/csrc/philox.cuh (TEST ONLY).

Philox4x32-10 (Salmon et al., SC'11) with the same counter layout and draw
conversions as the CUDA kernels, vectorised over environments, so the oracle
can reproduce the native-noise trajectories: integer words and the 53-bit
uniforms bit for bit, normals to libm accuracy (device log/cospi vs numpy).
"""
import numpy as np

STREAM_STEP, STREAM_RESET, STREAM_ACTION, STREAM_IMAGE = 0, 1, 2, 3
STREAM_NORMAL, STREAM_AUTORESET = 4, 5
STREAM_ZIG, STREAM_ZIG_RETRY = 6, 0x100
STREAM_GRID_ZIG, STREAM_ZIG_DRAW_RETRY = 44, 0x1000
STREAM_STATE_NOISE, STREAM_RESET_BOX = 8, 64
STREAM_IRR_STEP, STREAM_IRR_AUTORESET = 32, 33

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed):
    """Counters broadcast to a common shape; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = (c.astype(np.uint64) for c in (c0, c1, c2, c3))
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0
            p1 = _M1 * c2
            n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
            n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
            c1 = p1 & _LO
            c3 = p0 & _LO
            c0, c2 = n0 & _LO, n2 & _LO
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def uniform53(lo, hi):
    x = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform32(w):
    return (w.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def normal_pair_f64(a, b):
    """Box-Muller pair (philox.cuh normal_pair_f64)."""
    r = np.sqrt(-2.0 * np.log(uniform32(a)))
    ang = 2.0 * np.pi * uniform32(b)
    return r * np.cos(ang), r * np.sin(ang)


def normal_pair_fast(a, b):
    """fp32 restatement of philox.cuh normal_pair_fast (the device uses SFU
    approximations, so agreement is ~1e-6 relative, not bitwise)."""
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) \
        * np.float32(1.0 / 16777216.0)
    u2 = ((b >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) \
        * np.float32(1.0 / 16777216.0)
    r = np.sqrt(np.float32(-1.3862943611198906) * np.log2(u1))
    ang = np.float32(6.283185307179586) * u2
    return ((r * np.cos(ang)).astype(np.float64),
            (r * np.sin(ang)).astype(np.float64))


def _quad_words(seed, env_ids, step, stream):
    quad = int(step) >> 2
    return philox4x32_10(env_ids, quad & 0xFFFFFFFF, (quad >> 32) & 0xFFFFFFFF,
                         stream, seed)


def transition_noise_params(p, n_states):
    """Closed-form Philox-mode transition noise (csrc/context.cu): the
    reference's noisy distribution (rl_toy_env.py:1606-1617) gives P[s,a]
    probability 1 - p and each of the S-1 other states p / (S-1).  A 32-bit
    word w is noisy iff w < T = round(p 2^32); then floor(w M / 2^sh) with
    M = floor((S-1) 2^sh / T) < 2^32 (largest such sh) is uniform on [0, S-2]
    and the noisy state is (P[s,a] + 1 + k) mod S -- one of the S-1 others."""
    T = min(max(int(np.floor(float(p) * 4294967296.0 + 0.5)), 0), 1 << 32)
    if T == 0 or n_states < 2:
        return 0, 0, 32, n_states
    sh = 63
    while sh > 32 and (((n_states - 1) << sh) // T) >> 32:
        sh -= 1
    # T < S-1 (p below ~S 2^-32) cannot be uniform over the others anyway
    return T, min(((n_states - 1) << sh) // T, 0xFFFFFFFF), sh, n_states


def noisy_next_state(w, nxt, params):
    """Philox-mode noisy transition: words `w`, noise-free next states `nxt`."""
    T, M, sh, S = params
    w = np.asarray(w).astype(np.uint64)
    nxt = np.asarray(nxt, dtype=np.int64)
    # (w * M) >> sh with sh >= 32, as the device does: mulhi, then shift
    k = (((w * np.uint64(M)) >> np.uint64(32)) >> np.uint64(sh - 32)).astype(np.int64)
    return np.where(w < np.uint64(T), (nxt + 1 + k) % S, nxt)


def ziggurat_normal(seed, env_ids, step):
    """N(0,1) of global step `step` by csrc/ziggurat.cuh's scheme: the first
    64-bit word is (w1:w0) / (w3:w2) of Philox(env, step >> 1, STREAM_ZIG) for
    even / odd steps; a rejected first attempt continues numpy's algorithm on
    the words (q_2c, q_2c+1) = Philox(env, step, STREAM_ZIG_RETRY + c)."""
    from . import ziggurat as zg
    step = int(step)
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    pair = step >> 1
    w = philox4x32_10(env_ids, pair & 0xFFFFFFFF, (pair >> 32) & 0xFFFFFFFF,
                      STREAM_ZIG, seed)
    lo, hi = (w[2], w[3]) if step & 1 else (w[0], w[1])
    main = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)

    def retry(i, k):
        q = philox4x32_10(env_ids[i:i + 1], step & 0xFFFFFFFF,
                          (step >> 32) & 0xFFFFFFFF, STREAM_ZIG_RETRY + (k >> 1), seed)
        a, b = (q[2], q[3]) if k & 1 else (q[0], q[1])
        return (int(b[0]) << 32) | int(a[0])
    return zg.standard_normal_counter(main, retry)


def ziggurat_draw(seed, env_ids, lo, hi, step, draw):
    """N(0,1) from the first 64-bit words hi:lo (uint32 arrays, one per env) by
    csrc/ziggurat.cuh's zig_resolve_draw: draw `draw` of global step `step` continues,
    after a rejected first attempt, on the words (q_2c, q_2c+1) =
    Philox(env, step, STREAM_ZIG_DRAW_RETRY + 64 draw + c)."""
    from . import ziggurat as zg
    step = int(step)
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    main = (np.asarray(hi).astype(np.uint64) << np.uint64(32)) | np.asarray(lo).astype(np.uint64)

    def retry(i, k):
        q = philox4x32_10(env_ids[i:i + 1], step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF,
                          STREAM_ZIG_DRAW_RETRY + 64 * int(draw) + (k >> 1), seed)
        a, b = (q[2], q[3]) if k & 1 else (q[0], q[1])
        return (int(b[0]) << 32) | int(a[0])
    return zg.standard_normal_counter(main, retry)


def step_noise(seed, env_ids, step, want_normal=True, fast=False, raw=False,
               normal=None):
    """(transition uniform, N(0,1)) of global step `step` (`raw`: the 32-bit
    transition word instead of the uniform).  The draws come in
    groups of 4 steps (counter = step >> 2): STREAM_STEP word j is the 32-bit
    transition uniform of step 4q+j; STREAM_NORMAL words (0,1) and (2,3) feed
    two Box-Muller pairs = the 4 reward normals."""
    j = int(step) & 3
    u = _quad_words(seed, env_ids, step, STREAM_STEP)[j]
    if not raw:
        u = uniform32(u)
    z = None
    if want_normal and normal == "ziggurat":
        z = ziggurat_normal(seed, env_ids, step)
    elif want_normal:
        w = _quad_words(seed, env_ids, step, STREAM_NORMAL)
        pair = normal_pair_fast if fast else normal_pair_f64
        z = pair(w[0], w[1])[j] if j < 2 else pair(w[2], w[3])[j - 2]
    return u, z


def autoreset_uniform(seed, env_ids, step):
    """32-bit uniform of the same-step auto-reset after global step `step`."""
    return uniform32(_quad_words(seed, env_ids, step, STREAM_AUTORESET)[int(step) & 3])


def quad_word(seed, env_ids, step, stream):
    """Word (step & 3) of the 4-step group `step >> 2` of a per-step stream."""
    return _quad_words(seed, env_ids, step, stream)[int(step) & 3]


def step_words(seed, env_ids, step, stream=STREAM_STEP):
    step = int(step)
    return philox4x32_10(env_ids, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF,
                         stream, seed)


def mulhi32(w, n):
    return ((w.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
