"""numpy model of the bounded uint8 pipeline in porespy_b200/csrc/lt_kernels.cuh
(classify -> x-distance -> windowed 2-D distance -> reach -> cone scan).  Used by the CPU suite
to check the *derivation* (reach formula, cone scan, background skipping, caps) against the
oracle; the CUDA code itself is checked on the GPU."""
import numpy as np

CLS_BG, CLS_NEVER, GX_BG, GX_FAR = 255, 254, 255, 254


def isqrt(v):
    r = int(np.sqrt(float(v)))
    while r * r > v:
        r -= 1
    while (r + 1) * (r + 1) <= v:
        r += 1
    return r


def classify(d2, T):
    T = np.asarray(T, dtype=np.int64)
    d = d2.astype(np.int64)
    # first k with T[k] <= d  (T strictly descending) == number of thresholds > d
    k = (T[None, :] > d.reshape(-1, 1)).sum(axis=1).reshape(d2.shape)
    cls = np.where(k == len(T), CLS_NEVER, k)
    return np.where(d == 0, CLS_BG, cls).astype(np.uint8)


def x_distance(seeds):
    """distance along the last axis to the nearest True, capped at GX_FAR"""
    n = seeds.shape[-1]
    pos = np.arange(n)
    last = np.where(seeds, pos, -10**6)
    last = np.maximum.accumulate(last, axis=-1)
    nxt = np.where(seeds, pos, 10**6)
    nxt = np.minimum.accumulate(nxt[..., ::-1], axis=-1)[..., ::-1]
    d = np.minimum(pos - last, nxt - pos)
    return np.minimum(d, GX_FAR)


def cone_fill(m):
    """fill(z) <=> exists z': |z-z'| < m(z'), via the two sweeps c = max(m, c-1) along axis 0"""
    m = m.astype(np.int64)
    cf = np.zeros_like(m)
    c = np.zeros(m.shape[1:], dtype=np.int64)
    for z in range(m.shape[0]):
        c = np.maximum(m[z], c - 1)
        cf[z] = c
    cb = np.zeros_like(m)
    c = np.zeros(m.shape[1:], dtype=np.int64)
    for z in range(m.shape[0] - 1, -1, -1):
        c = np.maximum(m[z], c - 1)
        cb[z] = c
    return (cf > 0) | (cb > 0)


def reach_map(cls, k, T):
    """the xy kernel: uint8 reach m = ceil(sqrt(T - h)) (0 where h >= T or background)"""
    nz, ny, nx = cls.shape
    W = isqrt(T - 1)
    assert W <= 253
    gx = x_distance(cls <= k).astype(np.int64)
    bg = cls == CLS_BG
    gx = np.where(bg, GX_BG, gx)
    best = np.minimum(T, gx * gx)
    pad = np.full((nz, W, nx), GX_FAR, dtype=np.int64)
    g = np.concatenate([pad, gx, pad], axis=1)
    for dy in range(1, W + 1):
        up = g[:, W - dy:W - dy + ny, :]
        dn = g[:, W + dy:W + dy + ny, :]
        best = np.minimum(best, np.minimum(up * up, dn * dn) + dy * dy)
    m = np.where(bg | (best >= T), 0, np.ceil(np.sqrt((T - best).clip(min=0).astype(np.float32))))
    return m.astype(np.uint8)


def lt_idx(d2, T_list, seeds_filter=None):
    """uint8 radius-index map; seeds_filter(k, seeds) -> trimmed seeds (access-limited)."""
    d2 = d2.reshape((1,) * (3 - d2.ndim) + d2.shape)
    cls = classify(d2, T_list)
    idx = np.zeros(d2.shape, dtype=np.uint8)
    rcls = np.where(cls == CLS_BG, CLS_BG, CLS_NEVER).astype(np.uint8)
    for k, T in enumerate(T_list):
        cmap = cls
        if seeds_filter is not None:
            keep = seeds_filter(k, cls <= k)
            rcls = np.where(keep & (rcls == CLS_NEVER), k, rcls).astype(np.uint8)
            cmap = rcls
        fill = cone_fill(reach_map(cmap, k, int(T)))
        idx[(idx == 0) & fill] = k + 1
    return idx
