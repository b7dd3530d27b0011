"""Second, independent restatement of a few reference kernels in pure Python with np.float32
scalars (slow; tiny inputs only).  Used to cross-check the C oracle — test infrastructure."""
import math

import numpy as np

f32 = np.float32


def roi_align_v2_element(data, rois, N, PH, PW, scale, n, c, ph, pw):
    """roi_align_v2-inl.h:61-153 for one output element -> (val, argx, argy)."""
    _, C, H, W = data.shape
    r = rois.reshape(-1, 4)[n]
    b = n // N
    scale = f32(scale)
    rsw, rsh, rew, reh = (f32(r[0]) * scale, f32(r[1]) * scale, f32(r[2]) * scale, f32(r[3]) * scale)
    bh = f32(reh - rsh) / f32(PH)
    bw = f32(rew - rsw) / f32(PW)

    def clip(v, lim):
        v = v if v > f32(0) else f32(0)
        return v if v < f32(lim) else f32(lim)

    hs = clip(f32(f32(ph) * bh) + rsh, H - 1)
    he = clip(f32(f32(ph + 1) * bh) + rsh, H - 1)
    ws = clip(f32(f32(pw) * bw) + rsw, W - 1)
    we = clip(f32(f32(pw + 1) * bw) + rsw, W - 1)
    if he <= hs or we <= ws:
        return f32(0), f32(-1), f32(-1)
    best, bx, by = f32(-np.finfo(np.float32).max), f32(-1), f32(-1)
    hst = f32(float(f32(he - hs)) / 3.0)
    wst = f32(float(f32(we - ws)) / 3.0)
    plane = data[b, c]
    h = f32(hs + hst)
    while float(h) <= float(f32(he - hst)) + 0.01:
        w = f32(ws + wst)
        while float(w) <= float(f32(we - wst)) + 0.01:
            hl = min(max(int(math.floor(h)), 0), H - 1)
            hh = min(max(int(math.ceil(h)), 0), H - 1)
            wl = min(max(int(math.floor(w)), 0), W - 1)
            wr = min(max(int(math.ceil(w)), 0), W - 1)
            al = f32(0.5) if hl == hh else f32(h - f32(hl)) / f32(hh - hl)
            be = f32(0.5) if wl == wr else f32(w - f32(wl)) / f32(wr - wl)
            one = f32(1)
            v = f32(f32(f32(f32(one - al) * f32(one - be)) * plane[hl, wl]
                        + f32(f32(al * f32(one - be)) * plane[hh, wl]))
                    + f32(f32(f32(one - al) * be) * plane[hl, wr]))
            v = f32(v + f32(f32(al * be) * plane[hh, wr]))
            if v > best:
                best, bx, by = v, w, h
            w = f32(w + (wst if wst > f32(0.01) else f32(0.01)))
        h = f32(h + (hst if hst > f32(0.01) else f32(0.01)))
    return best, bx, by
