// flood_kernels.cuh -- inlet-connected flooding as lock-free union-find label propagation.
//
// Replaces trim_disconnected_blobs (/root/reference/src/porespy/filters/_funcs.py:1252-1270:
// scipy.ndimage.label of (inlets | im) + np.isin of the labels that touch an inlet).
//
// Nodes: voxel v is node v+1; node 0 is a virtual root that every inlet voxel points to from
// the start, so "connected to an inlet" <=> find(v+1) == 0.  Links always point from the
// larger to the smaller node id (atomicMin), hence node 0 is the root of its component and
// the structure is a forest at all times.  Inlet voxels are graph nodes even where the image
// is solid (F:1265: label(inlets + (im > 0))).
//
// In the porosimetry loop the seed sets are nested (class <= k grows with k), so the forest
// is kept across radii: each radius only links the voxels that became seeds since the
// previous radius (uf_activate) and then re-tests the not-yet-reached seeds (uf_mark).
#pragma once
#include "common.cuh"

struct InletSpec {
    int mode;                 // 1 = faces predicate, 2 = mask
    int ndim;                 // dimensionality of the squeezed image (faces predicate)
    const uint8_t *mask;      // mode 2 (the local slab of the mask)
    int z0, nzg;              // z-slab shard: local plane z is global plane z + z0 of nzg planes
};

__device__ __forceinline__ bool is_inlet(const InletSpec &s, int64_t v, int z, int y, int x,
                                         int nz, int ny, int nx)
{
    if (s.mode == 2) return s.mask[v] != 0;
    // get_border(shape, mode='faces') (generators/_borders.py:93-100); ndim 1: all True
    if (s.ndim >= 3) return z + s.z0 == 0 || z + s.z0 == s.nzg - 1 || y == 0 || y == ny - 1 || x == 0 || x == nx - 1;
    if (s.ndim == 2) return y == 0 || y == ny - 1 || x == 0 || x == nx - 1;
    return true;
}

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
{
    volatile uint32_t *vp = parent;
    uint32_t p = vp[x];
    while (p != x) {
        const uint32_t gp = vp[p];
        if (gp == p) return p;         // p is the root
        vp[x] = gp;                    // path halving (benign race: gp is always an ancestor, gp < x)
        x = gp;
        p = vp[x];
    }
    return x;
}

__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b)
{
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }     // a > b: hang a under b
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;                                   // a was still a root: linked
        a = old;                                                // lost a race: retry from its new parent
    }
}

__global__ void __launch_bounds__(256)
uf_init_kernel(uint32_t *__restrict__ parent, InletSpec inl, int nz, int ny, int nx)
{
    const int64_t n = (int64_t)nz * ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const int x = (int)(v % nx);
        const int64_t t = v / nx;
        const int y = (int)(t % ny), z = (int)(t / ny);
        parent[v + 1] = is_inlet(inl, v, z, y, x, nz, ny, nx) ? 0u : (uint32_t)(v + 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) parent[0] = 0u;
}

// Link every voxel with klo < cls <= khi to its active neighbours (active: inlet or cls <= khi).
// conn: 6 (faces) or 26 (faces+edges+corners); with nz == 1 these are 4- and 8-connectivity.
__global__ void __launch_bounds__(256)
uf_activate_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int klo,
                   int khi, int conn, int nz, int ny, int nx)
{
    const int64_t n = (int64_t)nz * ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const int c = cls[v];
        if (c <= klo || c > khi) continue;
        const int x = (int)(v % nx);
        const int64_t t = v / nx;
        const int y = (int)(t % ny), z = (int)(t / ny);
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int nn = (dz != 0) + (dy != 0) + (dx != 0);
                    if (nn == 0 || (conn == 6 && nn != 1)) continue;
                    const int zz = z + dz, yy = y + dy, xx = x + dx;
                    if (zz < 0 || zz >= nz || yy < 0 || yy >= ny || xx < 0 || xx >= nx) continue;
                    const int64_t u = ((int64_t)zz * ny + yy) * nx + xx;
                    if ((int)cls[u] <= khi || is_inlet(inl, u, zz, yy, xx, nz, ny, nx))
                        uf_union(parent, (uint32_t)(v + 1), (uint32_t)(u + 1));
                }
    }
}

// rcls[v] = k for every seed (cls <= k) that is connected to the inlets and was not marked
// at an earlier radius; *gate (monotone) is set once any voxel has ever been marked.
__global__ void __launch_bounds__(256)
uf_mark_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, uint8_t *__restrict__ rcls,
               int k, int64_t n, int *gate)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    int any = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        if ((int)cls[v] > k || rcls[v] != CLS_NEVER) continue;
        const uint32_t root = uf_find(parent, (uint32_t)(v + 1));
        if (root == 0u) { rcls[v] = (uint8_t)k; any = 1; }
        else ((volatile uint32_t *)parent)[v + 1] = root;        // compress
    }
    if (__any_sync(0xFFFFFFFFu, any) && lane_id() == 0 && *((volatile int *)gate) == 0) *gate = 1;
}

// ---- z-slab shards: the union-find is slab-local; connectivity through a slab face travels as
// one byte per face voxel ("this node is connected to the inlets") and is injected on the
// other side as a link to the virtual root.  Repeated by the host until no rank changes.
// out[i] = 1 if voxel i of local plane z is a graph node (inlet or cls <= k) whose root is 0.
__global__ void __launch_bounds__(256)
uf_face_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int k, int z,
               int nz, int ny, int nx, uint8_t *__restrict__ out)
{
    const int64_t plane = (int64_t)ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += step) {
        const int64_t v = (int64_t)z * plane + i;
        const int y = (int)(i / nx), x = (int)(i % nx);
        const bool node = (int)cls[v] <= k || is_inlet(inl, v, z, y, x, nz, ny, nx);
        out[i] = (node && uf_find(parent, (uint32_t)(v + 1)) == 0u) ? 1 : 0;
    }
}

// nb[i] != 0: the 6-neighbour of voxel i of local plane z across the slab face is connected to
// the inlets.  Every node of the plane under such a neighbour is linked to the root; *changed
// is set when that reached a component that was not connected before.
__global__ void __launch_bounds__(256)
uf_inject_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int k, int z,
                 int nz, int ny, int nx, const uint8_t *__restrict__ nb, int *changed)
{
    const int64_t plane = (int64_t)ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += step) {
        if (nb[i] == 0) continue;
        const int64_t v = (int64_t)z * plane + i;
        const int y = (int)(i / nx), x = (int)(i % nx);
        if (!((int)cls[v] <= k || is_inlet(inl, v, z, y, x, nz, ny, nx))) continue;
        if (uf_find(parent, (uint32_t)(v + 1)) == 0u) continue;
        uf_union(parent, (uint32_t)(v + 1), 0u);
        *changed = 1;
    }
}

// rcls init: background stays background, every foreground voxel is "not reached yet".
__global__ void __launch_bounds__(256)
uf_rcls_init_kernel(const uint8_t *__restrict__ cls, uint8_t *__restrict__ rcls, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        rcls[i] = cls[i] == CLS_BG ? CLS_BG : CLS_NEVER;
}

// standalone flood: class map of a binary mask (0 = foreground node, 255 = not in mask)
__global__ void __launch_bounds__(256)
flood_cls_kernel(const uint8_t *__restrict__ mask, uint8_t *__restrict__ cls, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        cls[i] = mask[i] ? 0 : CLS_BG;
}

__global__ void __launch_bounds__(256)
flood_out_kernel(const uint8_t *__restrict__ rcls, uint8_t *__restrict__ out, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        out[i] = rcls[i] == 0 ? 1 : 0;
}
