"""numpy restatement of the CUDA renderers' algorithm (TEST INFRASTRUCTURE).

Lets the CPU suite validate the host-built image tables
(mdp_playground_b200/image_tables.py) and the gather logic of
csrc/render_kernels.cuh against Pillow without a GPU: same mask atlas, same
16.16 affine gather, same flip / transpose order, vectorised over pixels."""
import numpy as np

from mdp_playground_b200 import image_tables as it


def render_discrete(tb, state, R, shift_w, shift_h, rotation, flip):
    """uint8 [W][H] (obs[x][y] = pil[y][x]) of one polygon image."""
    W, H = tb.width, tb.height
    ri = R - tb.r_min
    mid = tb.mask_index[state, ri, tb.xvar[state, ri, shift_w],
                        tb.yvar[state, ri, shift_h]]
    mask = tb.mask_bits[mid]
    ys, xs = np.mgrid[0:H, 0:W].astype(np.int64)   # final PIL image coords
    fx, fy = xs, ys
    if flip == 1:
        fx = W - 1 - xs
    elif flip == 2:
        fy = H - 1 - ys
    valid = np.ones((H, W), dtype=bool)
    if rotation is not None and rotation >= 0:
        a0, a1, a2, a3, a4, a5 = (int(v) for v in tb.rot_coeff[rotation % 360])
        rx = (a2 + a1 * fy + a0 * fx) >> 16
        ry = (a5 + a4 * fy + a3 * fx) >> 16
        valid = (rx >= 0) & (rx < W) & (ry >= 0) & (ry < H)
    else:
        rx, ry = fx, fy
    mx = rx - shift_w + it.MASK_CENTRE
    my = ry - shift_h + it.MASK_CENTRE
    inside = valid & (mx >= 0) & (mx < it.MASK_ROWS) & (my >= 0) & (my < it.MASK_ROWS)
    cols = mask[np.clip(mx, 0, it.MASK_ROWS - 1)]   # column-major bitmaps
    bit = (cols >> np.clip(my, 0, 63).astype(np.uint64)) & np.uint64(1)
    pil = np.where(inside & (bit == 1), 255, 0).astype(np.uint8)  # [y][x]
    return pil.T.copy()


def radius_from_uniform(tb, u):
    return tb.r_min + int(np.searchsorted(tb.r_thresholds, u, side="right"))
