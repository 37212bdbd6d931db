"""Sequential Python model of the flood's record rules (porespy_b200/csrc/flood_kernels.cuh, DESIGN.md lemma x):
row chains pre-linked "downhill", x records where no chain covers an x edge, records of the other directions at the
local minima of the pair activation index, unions in index order with join times, one resolve.  Test infrastructure:
it states the rules the CUDA kernels implement in a form that can be checked against plain labelling on the CPU."""
import numpy as np

FAM = [(0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 1, -1), (0, 1, 1), (1, 0, -1), (1, 0, 1), (1, 1, -1), (1, 1, 0), (1, 1, 1),
       (1, -1, -1), (1, -1, 0), (1, -1, 1)]


def reached_class(cls, inlet, conn=6, seg=128, inlets_in_set=False):
    """rcls[v] = first index at which voxel v (class cls[v] < 254) is a node connected to the inlets; 254 never, 255
    where cls is 255.  `inlets_in_set`: an inlet voxel only counts from its own class on (and keeps its class)."""
    nz, ny, nx = cls.shape
    n = cls.size
    if inlets_in_set:
        inlet = inlet & (cls < 254)
        a = cls.astype(np.int64)
    else:
        a = np.where(inlet, 0, cls).astype(np.int64)
    parent = np.arange(n + 1)
    jt = np.full(n + 1, 255)

    def vid(z, y, x):
        return (z * ny + y) * nx + x

    for z in range(nz):                                   # uf_prelink_kernel
        for y in range(ny):
            for x0 in range(0, nx, seg):
                ln = min(seg, nx - x0)
                p = list(range(ln))
                for i in range(ln):
                    ac = a[z, y, x0 + i]
                    if inlet[z, y, x0 + i] or ac >= 254:
                        continue
                    if i > 0 and a[z, y, x0 + i - 1] <= ac:
                        p[i] = i - 1
                    elif i + 1 < ln and a[z, y, x0 + i + 1] < ac:
                        p[i] = i + 1
                for i in range(ln):
                    r = i
                    while p[r] != r:
                        r = p[r]
                    hang = inlet[z, y, x0 + i] or inlet[z, y, x0 + r]
                    parent[vid(z, y, x0 + i) + 1] = 0 if hang else vid(z, y, x0 + r) + 1
                    if hang:
                        jt[vid(z, y, x0 + i) + 1] = 0
    nfam = 3 if conn == 6 else 13
    recs = {}

    def A(z, y, x):
        return a[z, y, x] if 0 <= z < nz and 0 <= y < ny and 0 <= x < nx else 255

    for z in range(nz):                                   # uf_emit_kernel
        for y in range(ny):
            for x in range(nx):
                ax = a[z, y, x]
                if ax >= 254:
                    continue
                x0 = (x // seg) * seg
                ln, i = min(seg, nx - x0), x - x0
                ar = A(z, y, x + 1)
                if ar < 254:
                    inx, inr = bool(inlet[z, y, x]), bool(inlet[z, y, x + 1])
                    cov = inx and inr
                    if not cov and i + 1 < ln:
                        left_ok = i > 0 and A(z, y, x - 1) <= ax
                        cov = ((not inr) and ax <= ar) or ((not inx) and (not left_ok) and ar < ax)
                    if not cov:
                        recs.setdefault(max(ax, ar), []).append((vid(z, y, x), 0))
                for f in range(1, nfam):
                    dz, dy, dx = FAM[f]
                    bx = A(z + dz, y + dy, x + dx)
                    if bx >= 254:
                        continue
                    m = max(ax, bx)
                    ml = max(A(z, y, x - 1), A(z + dz, y + dy, x - 1 + dx))
                    mr = max(A(z, y, x + 1), A(z + dz, y + dy, x + 1 + dx))
                    if ml > m and mr >= m:
                        recs.setdefault(m, []).append((vid(z, y, x), f))

    def find(x):
        while x != 0 and parent[x] != x:
            x = parent[x]
        return x

    for k in sorted(recs):                                # uf_union_rec_kernel, one launch per index
        for v, f in recs[k]:
            dz, dy, dx = FAM[f]
            ra, rb = find(v + 1), find(v + (dz * ny + dy) * nx + dx + 1)
            if ra == rb:
                continue
            if ra < rb:
                ra, rb = rb, ra
            if rb == 0:
                jt[ra] = k
            parent[ra] = rb
    out = np.full(n, 254)                                 # uf_resolve_kernel
    flat = cls.ravel()
    for v in range(n):
        c = flat[v]
        if c == 255:
            out[v] = 255
        elif c < 254:
            x = v + 1
            while parent[x] != 0 and parent[x] != x:
                x = parent[x]
            if parent[x] == 0:
                out[v] = max(c, jt[x])
    return out.reshape(cls.shape), sum(len(r) for r in recs.values())


def reached_class_by_labelling(cls, inlet, conn=6, inlets_in_set=False):
    """The same map from one scipy labelling per index (what the reference does per radius / bin)."""
    import scipy.ndimage as spim
    st = spim.generate_binary_structure(3, 1 if conn == 6 else 3)
    out = np.where(cls == 255, 255, 254)
    top = int(cls[cls < 254].max(initial=-1))
    for k in range(top, -1, -1):
        nodes = (cls <= k) if inlets_in_set else ((cls <= k) | inlet)
        lab = spim.label(nodes, structure=st)[0]
        keep = np.unique(lab[inlet & nodes])
        out[np.isin(lab, keep[keep > 0]) & (cls <= k)] = k
    return out
