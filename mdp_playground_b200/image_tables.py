"""Host-built tables of the image-observation renderers.

The reference renders with Pillow (ImageDraw.polygon / ellipse / rectangle,
Image.rotate, Image.transpose; spaces/image_multi_discrete.py:129-270,
spaces/image_continuous.py:116-208).  Like the MDP tables these small,
finite tables are produced on the host -- by Pillow itself where its raster
rules matter -- and uploaded once; the per-step work (which mask, where, the
rotation gather, flip, transpose, 128-bit stores) runs on the GPU.

  * polygon mask atlas: one bit mask per (state, radius R, x-vertex variant,
    y-vertex variant).  Pillow's fill is translation invariant for identical
    integer vertex offsets, but the reference's vertices
    `int(shift + R*cos(a))` (:188-194) are fp64-rounded per shift, so each
    (state, R) has up to a few distinct offset tuples ("variants") selected
    by the actual shift (SURVEY.md finding 4).
  * rotation: Image.rotate(NEAREST) is a 16.16 fixed-point inverse affine
    gather; the 6 integer coefficients per angle follow Pillow's own
    arithmetic and are self-checked against Pillow at build time.
  * ImageContinuous: the 11x11 ellipse stamp of radius 5.
"""
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

MASK_ROWS = 64           # rows of one mask, each a 64-bit word
MASK_CENTRE = 31         # mask pixel of the polygon centre (shift_w, shift_h)
CIRCLE_RADIUS = 20       # rl_toy_env.py:715
MAX_RADIUS = MASK_CENTRE - 1


def _fix(v):
    return int(math.floor(v * 65536.0 + 0.5))


def rotation_coefficients(width, height, check=True):
    """int32[360][6]: (a0..a5) of xin = (a2 + a1*y + a0*x) >> 16,
    yin = (a5 + a4*y + a3*x) >> 16 for Image.rotate(angle) (NEAREST, no
    expand, centre (w/2, h/2)), including Pillow's transpose fast paths for
    0 / 180 / (90, 270 when square) expressed as exact integer affines."""
    w, h = width, height
    one = 65536
    out = np.zeros((360, 6), dtype=np.int64)
    for deg in range(360):
        if deg == 0:
            c = [one, 0, 0, 0, one, 0]
        elif deg == 180:  # ROTATE_180: out(x, y) = in(w-1-x, h-1-y)
            c = [-one, 0, (w - 1) * one, 0, -one, (h - 1) * one]
        elif deg == 90 and w == h:  # ROTATE_90: out(x, y) = in(w-1-y, x)
            c = [0, -one, (w - 1) * one, one, 0, 0]
        elif deg == 270 and w == h:  # ROTATE_270: out(x, y) = in(y, h-1-x)
            c = [0, one, 0, -one, 0, (h - 1) * one]
        else:
            ang = -math.radians(deg)
            m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0,
                 round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
            cx, cy = w / 2, h / 2
            m[2] = m[0] * -cx + m[1] * -cy + m[2]
            m[5] = m[3] * -cx + m[4] * -cy + m[5]
            m[2] += cx
            m[5] += cy
            c = [_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
                 _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5)]
        out[deg] = c
    assert np.abs(out).max() < 2**31
    out = out.astype(np.int32)
    if check:
        _check_rotation(out, w, h)
    return out


def apply_rotation_map(coeff, width, height):
    """(xin, yin, valid) int arrays [H][W] of one angle's gather."""
    ys, xs = np.mgrid[0:height, 0:width].astype(np.int64)
    a0, a1, a2, a3, a4, a5 = (int(v) for v in coeff)
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    valid = (xin >= 0) & (xin < width) & (yin >= 0) & (yin < height)
    return xin, yin, valid


def _check_rotation(coeffs, w, h):
    """Pillow is the authority: rotate an index image ('I', int32) by every
    angle and compare with the integer gather."""
    from PIL import Image
    idx = np.arange(1, w * h + 1, dtype=np.int32).reshape(h, w)
    img = Image.fromarray(idx, mode="I")
    for deg in range(360):
        want = np.array(img.rotate(deg))
        xin, yin, valid = apply_rotation_map(coeffs[deg], w, h)
        got = np.where(valid, idx[np.clip(yin, 0, h - 1), np.clip(xin, 0, w - 1)], 0)
        if not np.array_equal(got, want):
            raise RuntimeError(
                f"rotation table disagrees with Pillow at {deg} degrees for "
                f"{w}x{h}: this Pillow build uses a different affine path")


@dataclass
class DiscreteImageTables:
    width: int
    height: int
    n_states: int
    r_min: int
    n_radii: int
    n_xvar: int
    n_yvar: int
    has_scale: bool
    has_shift: bool
    has_rotate: bool
    has_flip: bool
    sh_quant: int
    ro_quant: int
    mask_bits: np.ndarray      # uint64 [n_masks][MASK_ROWS], column bitmaps
    mask_index: np.ndarray     # int32 [S][n_radii][n_xvar][n_yvar]
    xvar: np.ndarray           # uint8 [S][n_radii][W]   variant of shift_w
    yvar: np.ndarray           # uint8 [S][n_radii][H]   variant of shift_h
    rot_coeff: np.ndarray      # int32 [360][6]
    r_thresholds: np.ndarray   # float64 [n_radii-1]  u-thresholds of R
    log_r_min: float
    log_r_span: float


def _vertex_offsets(sides, R, shift, fn):
    """Offsets of the reference's vertices from the shift (:186-194)."""
    return tuple(int(shift + R * fn((2 * np.pi / sides) * i)) - shift
                 for i in range(sides))


def _polygon_mask(dx, dy):
    """Pillow polygon fill of the vertex offsets around MASK_CENTRE."""
    from PIL import Image, ImageDraw
    img = Image.new("L", (MASK_ROWS, MASK_ROWS))
    pts = [(MASK_CENTRE + x, MASK_CENTRE + y) for x, y in zip(dx, dy)]
    ImageDraw.Draw(img).polygon(pts, fill=255)
    arr = np.array(img) > 0            # [y][x]
    # column-major bitmaps: word x holds the column, bit y is pixel (x, y).
    # The observation is the transposed image (obs[x][y] = pil[y][x]), so 16
    # consecutive output bytes are 16 consecutive bits of one word.
    bits = np.zeros(MASK_ROWS, dtype=np.uint64)
    for x in range(MASK_ROWS):
        v = 0
        for y in np.nonzero(arr[:, x])[0]:
            v |= 1 << int(y)
        bits[x] = v
    return bits


def build_discrete_image_tables(n_states, width, height, transforms, sh_quant,
                                ro_quant, scale_range) -> DiscreteImageTables:
    W, H = int(width), int(height)
    has_scale = "scale" in transforms
    has_shift = "shift" in transforms
    has_rotate = "rotate" in transforms
    has_flip = "flip" in transforms
    R0 = CIRCLE_RADIUS
    if has_scale:
        lo, hi = scale_range
        log_min, log_max = np.log(lo * R0), np.log(hi * R0)
        # R = int(exp(log_min + u (log_max - log_min))), u in [0, 1)
        r_min = int(np.exp(log_min))
        r_max = int(np.exp(log_min + np.nextafter(1.0, 0.0) * (log_max - log_min)))
    else:
        log_min, log_max = 0.0, 1.0
        r_min = r_max = R0
    if r_max > MAX_RADIUS:
        raise NotImplementedError(f"polygon radius up to {r_max} > {MAX_RADIUS}")
    n_radii = r_max - r_min + 1
    # u-threshold of each radius step: R >= k  <=>  u >= (ln k - log_min)/span,
    # evaluated with the reference's own expression so ties fall the same way
    span = log_max - log_min
    thresholds = np.zeros(max(n_radii - 1, 0))
    for i, k in enumerate(range(r_min + 1, r_max + 1)):
        # smallest u with int(exp(log_min + u*span)) >= k, found by bisection
        # on the actual expression
        lo_u, hi_u = 0.0, 1.0
        for _ in range(80):
            mid = (lo_u + hi_u) / 2
            if int(np.exp(log_min + mid * span)) >= k:
                hi_u = mid
            else:
                lo_u = mid
        thresholds[i] = hi_u

    masks, mask_ids = [], {}
    xvar = np.zeros((n_states, n_radii, W), dtype=np.uint8)
    yvar = np.zeros((n_states, n_radii, H), dtype=np.uint8)
    per_cell = {}
    n_xvar = n_yvar = 1
    for s in range(n_states):
        sides = s + 3
        for ri, R in enumerate(range(r_min, r_max + 1)):
            xs, ys = {}, {}
            for sw in range(W):
                t = _vertex_offsets(sides, R, sw, np.cos)
                xvar[s, ri, sw] = xs.setdefault(t, len(xs))
            for sh in range(H):
                t = _vertex_offsets(sides, R, sh, np.sin)
                yvar[s, ri, sh] = ys.setdefault(t, len(ys))
            per_cell[(s, ri)] = (list(xs), list(ys))
            n_xvar, n_yvar = max(n_xvar, len(xs)), max(n_yvar, len(ys))
    mask_index = np.zeros((n_states, n_radii, n_xvar, n_yvar), dtype=np.int32)
    for (s, ri), (xs, ys) in per_cell.items():
        for xi, dx in enumerate(xs):
            for yi, dy in enumerate(ys):
                key = (dx, dy)
                if key not in mask_ids:
                    mask_ids[key] = len(masks)
                    masks.append(_polygon_mask(dx, dy))
                mask_index[s, ri, xi, yi] = mask_ids[key]
    return DiscreteImageTables(
        width=W, height=H, n_states=n_states, r_min=r_min, n_radii=n_radii,
        n_xvar=n_xvar, n_yvar=n_yvar, has_scale=has_scale, has_shift=has_shift,
        has_rotate=has_rotate, has_flip=has_flip,
        sh_quant=int(sh_quant or 1), ro_quant=int(ro_quant or 1),
        mask_bits=np.ascontiguousarray(np.stack(masks)),
        mask_index=mask_index, xvar=xvar, yvar=yvar,
        rot_coeff=rotation_coefficients(W, H) if has_rotate
        else np.zeros((360, 6), dtype=np.int32),
        r_thresholds=thresholds, log_r_min=float(log_min),
        log_r_span=float(span))


def disc_stamp(radius=5):
    """Row spans (x offset from the centre, width) of Pillow's filled ellipse
    with bounding box [c - r, c + r] (image_continuous.py:182-195)."""
    from PIL import Image, ImageDraw
    n = 4 * radius + 1
    c = 2 * radius
    img = Image.new("L", (n, n))
    ImageDraw.Draw(img).ellipse([(c - radius, c - radius),
                                 (c + radius, c + radius)], fill=255)
    arr = np.array(img) > 0
    spans = np.zeros((2 * radius + 1, 2), dtype=np.int32)
    for i, y in enumerate(range(c - radius, c + radius + 1)):
        xs = np.nonzero(arr[y])[0]
        assert len(xs) and np.all(np.diff(xs) == 1)
        spans[i] = (xs[0] - c, len(xs))
    assert not arr[:c - radius].any() and not arr[c + radius + 1:].any()
    return spans
