"""Independent vectorised numpy restatements of the integer stages (second opinion on the C++ oracle).

Arrays are [row, col].  Signed division truncates toward zero like Rust (numpy // floors)."""
import numpy as np


def tdiv(a, d):
    a = np.asarray(a, np.int32)
    return (np.sign(a) * (np.abs(a) // d)).astype(np.int32)


def mean_pyramid(img, max_levels):
    """src/core/multires.rs:21-31, 38-88."""
    out = [np.asarray(img, np.uint8)]
    while len(out) < max_levels:
        m = out[-1]
        hr, hc = m.shape[0] // 2, m.shape[1] // 2
        if hr == 0 or hc == 0:
            break
        m = m[:2 * hr, :2 * hc].astype(np.uint16)
        s = m[0::2, 0::2] + m[1::2, 0::2] + m[0::2, 1::2] + m[1::2, 1::2]
        out.append((s // 4).astype(np.uint8))
    return out


def centered(img):
    """src/core/gradient.rs:15-33."""
    m = np.asarray(img, np.int32)
    gx = np.zeros(m.shape, np.int16)
    gy = np.zeros(m.shape, np.int16)
    if m.shape[0] > 2 and m.shape[1] > 2:
        gx[1:-1, 1:-1] = tdiv(m[1:-1, 2:] - m[1:-1, :-2], 2)
        gy[1:-1, 1:-1] = tdiv(m[2:, 1:-1] - m[:-2, 1:-1], 2)
    return gx, gy


def bloc_gradients(finer):
    """src/core/gradient.rs:74-93 through multires::halve: a=(2i,2j) b=(2i+1,2j) c=(2i,2j+1) d=(2i+1,2j+1)."""
    m = np.asarray(finer, np.int32)
    hr, hc = m.shape[0] // 2, m.shape[1] // 2
    m = m[:2 * hr, :2 * hc]
    a, b, c, d = m[0::2, 0::2], m[1::2, 0::2], m[0::2, 1::2], m[1::2, 1::2]
    return tdiv(c + d - a - b, 2).astype(np.int16), tdiv(b - a + d - c, 2).astype(np.int16)


def gradients_tracker(pyr):
    """inverse_compositional.rs:112-117."""
    gxs, gys = [], []
    gx, gy = centered(pyr[0])
    gxs.append(gx)
    gys.append(gy)
    for l in range(1, len(pyr)):
        gx, gy = bloc_gradients(pyr[l - 1])
        gxs.append(gx)
        gys.append(gy)
    g2 = [(x.astype(np.int32) ** 2 + y.astype(np.int32) ** 2).astype(np.uint16) for x, y in zip(gxs, gys)]
    return gxs, gys, g2


def c2f_select(thresh, g2_levels):
    """src/core/candidates/coarse_to_fine.rs:15-89; g2 finest first -> masks finest first."""
    L = len(g2_levels)
    masks = [None] * L
    masks[L - 1] = np.ones(g2_levels[L - 1].shape, bool)
    for l in range(L - 2, -1, -1):
        g = np.asarray(g2_levels[l], np.uint16)
        pre = masks[l + 1]
        hr, hc = pre.shape
        assert (hr, hc) == (g.shape[0] // 2, g.shape[1] // 2)
        gg = g[:2 * hr, :2 * hc]
        vals = np.stack([gg[0::2, 0::2], gg[1::2, 0::2], gg[0::2, 1::2], gg[1::2, 1::2]], 0).reshape(4, -1)
        order = np.argsort(vals, axis=0, kind="stable")
        srt = np.take_along_axis(vals, order, 0)
        first, second = order[3], order[2]
        keep2 = srt[2] > (srt[1] + np.uint16(thresh)).astype(np.uint16)
        sel = np.zeros((4, vals.shape[1]), bool)
        cols = np.arange(vals.shape[1])
        sel[first, cols] = True
        sel[second[keep2], cols[keep2]] = True
        sel &= pre.reshape(1, -1)
        sel = sel.reshape(4, hr, hc)
        m = np.zeros(g.shape, bool)
        m[0:2 * hr:2, 0:2 * hc:2] = sel[0]
        m[1:2 * hr:2, 0:2 * hc:2] = sel[1]
        m[0:2 * hr:2, 1:2 * hc:2] = sel[2]
        m[1:2 * hr:2, 1:2 * hc:2] = sel[3]
        masks[l] = m
    return masks
