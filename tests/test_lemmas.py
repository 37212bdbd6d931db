"""CPU checks of the exactness lemmas the kernels rely on, independent of any GPU:

* lemma (v)  -- a min-plus result computed with values and offsets saturated at a cap is exact
  wherever it is below the cap (the 16-bit EDT passes, minplus_kernels.cuh);
* lemma (vi) -- in a union-find whose links are made root-to-root while the node sets grow, the
  link INTO the inlet root carries the step at which the subtree became connected, so
  `max(class, join time of the top node)` is the first radius at which a voxel is a reached seed
  (flood_kernels.cuh: skip rule for y/z links, star-linking of x-runs, path compression that carries
  the join time).  The model below follows the kernels' rules sequentially and is compared with the
  oracle's per-radius `trim_disconnected_blobs` (F:1181-1183, F:1252-1270).
"""
import numpy as np
import pytest

from oracle import cpu as oc

CAP = 32767


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_capped_minplus_is_exact_below_the_cap(seed):
    rng = np.random.default_rng(seed)
    n = 600
    f = rng.integers(33000, 60000, n).astype(np.int64)    # values above the cap ...
    f[rng.integers(0, n, 40)] = rng.integers(30000, 32767, 40)   # ... and some just below it, anywhere
    f[rng.random(n) < 0.3] = 2 ** 31                      # "no site in this line"
    f[rng.integers(0, 60, 6)] = rng.integers(0, 50, 6)    # a few near sites, all at one end of the line
    d = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :]).astype(np.int64)
    true = (f[None, :] + d * d).min(axis=1)
    capped = (np.minimum(f, CAP)[None, :] + np.minimum(d * d, CAP)).min(axis=1)
    below = true < CAP
    assert below.any() and (~below).any()
    assert np.array_equal(capped[below], true[below])      # exact wherever the result fits
    assert np.all(capped[~below] >= CAP)                   # and flagged (>= cap) everywhere else


# ------------------------------------------------------------------ lemma (vi): join times
UNSET = 255


class JoinTimeForest:
    """Sequential model of flood_kernels.cuh: node v + 1 per voxel, node 0 = the inlet root."""

    def __init__(self, acls):
        self.acls = acls                                   # activation level per voxel (inlets folded to 0)
        self.shape = acls.shape
        n = acls.size
        self.parent = np.arange(n + 1, dtype=np.int64)
        self.jtime = np.full(n + 1, UNSET, dtype=np.int64)

    def find(self, x):
        if x == 0:
            return 0
        p = self.parent[x]
        while p != x:
            if p == 0:
                return 0
            gp = self.parent[p]
            if gp == p:
                return p
            if gp == 0:                                    # p is a child of node 0: x joins node 0 with p's time
                if self.jtime[p] == UNSET:
                    return 0
                self.jtime[x] = self.jtime[p]
                self.parent[x] = 0
                return 0
            self.parent[x] = gp                            # path halving
            x = gp
            p = self.parent[x]
        return x

    def union(self, a, b, k):
        while True:
            a, b = self.find(a), self.find(b)
            if a == b:
                return
            if a < b:
                a, b = b, a
            if b == 0:
                self.jtime[a] = k
            old = self.parent[a]
            self.parent[a] = min(old, b)                   # atomicMin
            if old == a:
                return
            a = old

    def activate(self, k):
        nz, ny, nx = self.shape
        ac = self.acls.reshape(-1)
        new = np.flatnonzero(ac == k)
        isnew = lambda u: ac[u] == k
        active = lambda u: ac[u] <= k
        for v in new:                                      # every (voxel, direction) job of uf_union_list_kernel
            x, y, z = v % nx, (v // nx) % ny, v // (nx * ny)
            for dz, dy, dx in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
                zz, yy, xx = z + dz, y + dy, x + dx
                if not (0 <= zz < nz and 0 <= yy < ny and 0 <= xx < nx):
                    continue
                u = (zz * ny + yy) * nx + xx
                if isnew(u):
                    if u > v:
                        continue                           # both new: linked once, from the larger index
                elif not active(u):
                    continue
                if dx == 0 and x > 0 and active(v - 1) and active(u - 1):
                    continue                               # y / z pair with an active pair to its left
                w = u
                if dx == -1 and isnew(u):                  # link to the first voxel of the run of new voxels
                    xs, steps = xx, 0
                    while steps < 64 and xs > 0 and isnew(w - 1):
                        w, xs, steps = w - 1, xs - 1, steps + 1
                self.union(v + 1, w + 1, k)

    def resolve(self, cls):
        out = np.full(cls.size, 254, dtype=np.int64)
        flat = cls.reshape(-1)
        out[flat == 255] = 255
        for v in np.flatnonzero(flat < 254):
            x, p = v + 1, self.parent[v + 1]
            top = 0
            while True:
                if p == 0:
                    top = x
                    break
                if p == x:
                    break
                x, p = p, self.parent[p]
            if top:
                out[v] = max(flat[v], self.jtime[top])
        return out.reshape(cls.shape)


@pytest.mark.parametrize("shape,inlet", [((14, 13, 16), "z0"), ((12, 15, 14), "faces"), ((1, 30, 34), "x0")])
def test_join_times_give_the_first_reached_radius(shape, inlet):
    im = oc.blobs(list(shape), porosity=0.65, blobiness=0.9, seed=4) if shape[0] > 1 else \
        oc.blobs(list(shape[1:]), porosity=0.65, blobiness=1.2, seed=4).reshape(shape)
    d2 = oc.edt_sq(np.squeeze(im)).reshape(shape).astype(np.int64)
    T = [t for t in (26, 17, 10, 6, 4, 2, 1) if t <= d2.max()]
    cls = np.full(shape, 254, dtype=np.int64)
    cls[d2 == 0] = 255
    for k in reversed(range(len(T))):
        cls[d2 >= T[k]] = k                                # class = first (largest) threshold the voxel reaches
    inl = np.zeros(shape, dtype=bool)
    if inlet == "z0":
        inl[0] = True
    elif inlet == "x0":
        inl[:, :, 0] = True
    else:
        inl = oc.border_faces(np.squeeze(im).shape).reshape(shape)
    acls = np.where(inl, 0, cls)                           # inlet voxels are nodes from the first radius on
    forest = JoinTimeForest(acls)
    forest.parent[1:][inl.reshape(-1)] = 0
    forest.jtime[1:][inl.reshape(-1)] = 0
    for k in range(len(T)):
        forest.activate(k)
    got = forest.resolve(cls)
    # reference semantics, radius by radius: trim_disconnected_blobs with the cross strel (F:1181-1183)
    want = np.where(cls == 255, 255, 254)
    sq = np.squeeze(im)
    strel = oc._cross(sq.ndim)
    for k in reversed(range(len(T))):
        seeds = np.squeeze(d2 >= T[k])
        reached = oc.trim_disconnected_blobs(seeds, np.squeeze(inl), strel=strel).reshape(shape)
        want[reached.astype(bool)] = k
    assert np.array_equal(got, want)
    assert (got < 254).any()
