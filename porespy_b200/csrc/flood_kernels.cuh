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
// previous radius (uf_activate).
//
// Join times.  Links are made between roots, so along any path to the root the radius index at
// which each link was made never decreases, and the link INTO node 0 carries the index at which
// the whole subtree became connected to the inlets.  That index is kept in jtime[] for the
// children of node 0 (inlet voxels: 0), and path compression never bypasses a child of node 0.
// The single-GPU loop therefore runs all activations first and reads the first radius at which
// every voxel is a reached seed in ONE pass at the end (uf_resolve:  rcls = max(cls, jtime[top]))
// instead of re-testing the unreached seeds after every radius; the step-level ABI of the
// z-slab shards still marks per radius (uf_mark), because connectivity arrives through the slab
// faces between radii.
#pragma once
#include "common.cuh"

#define UF_TIME_UNSET 0xFFu   // jtime of a node that never hung directly under node 0

struct InletSpec {
    int mode;                 // 1 = faces predicate, 2 = mask, 3 = none (already folded into the class map),
                              // 4 = mask, but a voxel only acts as an inlet from its own class on (not a node before)
    int ndim;                 // dimensionality of the squeezed image (faces predicate)
    const uint8_t *mask;      // mode 2 (the local slab of the mask)
    int z0, nzg;              // z-slab shard: local plane z is global plane z + z0 of nzg planes
};

__device__ __forceinline__ bool is_inlet(const InletSpec &s, int64_t v, int z, int y, int x,
                                         int nz, int ny, int nx)
{
    if (s.mode == 2 || s.mode == 4) return s.mask[v] != 0;      // (mode 4: callers test the voxel's class as well)
    if (s.mode == 3) return false;          // inlets are folded into the class map (class 0), see uf_init_kernel
    // get_border(shape, mode='faces') (generators/_borders.py:93-100); ndim 1: all True
    if (s.ndim >= 3) return z + s.z0 == 0 || z + s.z0 == s.nzg - 1 || y == 0 || y == ny - 1 || x == 0 || x == nx - 1;
    if (s.ndim == 2) return y == 0 || y == ny - 1 || x == 0 || x == nx - 1;
    return true;
}

// Root of x, with path halving.  Node 0 is never read (it is the root of everything that reaches
// it): parent[0] and the few children of node 0 that carry whole percolating clusters would
// otherwise be fetched by every find of every launch -- one L2 slice serialising the kernel.
// With join times (jtime != NULL) a node is re-pointed from a child p of node 0 to node 0 itself
// only together with p's join time.  A time is written (once, or by several threads with the same
// value) just before the link that makes it meaningful; no fence orders the two, so a reader that
// sees the link but still the UF_TIME_UNSET sentinel simply does not re-point.
__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x, uint8_t *jtime = nullptr)
{
    if (x == 0u) return 0u;
    volatile uint32_t *vp = parent;
    uint32_t p = vp[x];
    while (p != x) {
        if (p == 0u) return 0u;
        const uint32_t gp = vp[p];
        if (gp == p) return p;         // p is the root
        if (gp == 0u) {                // p is a child of node 0: x joins node 0 directly
            if (jtime) {
                const uint8_t t = reinterpret_cast<volatile uint8_t *>(jtime)[p];
                if (t == UF_TIME_UNSET) return 0u;      // p's time is not visible yet: leave x where it is
                reinterpret_cast<volatile uint8_t *>(jtime)[x] = t;
            }
            vp[x] = 0u;
            return 0u;
        }
        vp[x] = gp;                    // path halving (benign race: gp is always an ancestor, gp < x)
        x = gp;
        p = vp[x];
    }
    return x;
}

// jtime != NULL: a node that is hung under node 0 records the radius index k of that event
// (written before the link, see uf_find).
__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b,
                                         uint8_t *jtime = nullptr, int k = 0)
{
    while (true) {
        a = uf_find(parent, a, jtime);
        b = uf_find(parent, b, jtime);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }     // a > b: hang a under b
        if (jtime && b == 0u) {
            // a was a root a moment ago, so if it hangs under node 0 already that happened in this
            // launch (same k); atomicMin lowers parent[a] to 0 whether or not a is still a root
            reinterpret_cast<volatile uint8_t *>(jtime)[a] = (uint8_t)k;
        }
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;                                   // a was still a root: linked
        a = old;                                                // lost a race: retry from its new parent
    }
}

__global__ void __launch_bounds__(256)
uf_init_kernel(uint32_t *__restrict__ parent, InletSpec inl, int nz, int ny, int nx, uint8_t *__restrict__ jtime,
               const uint8_t *__restrict__ cls = nullptr, uint8_t *__restrict__ acls = nullptr)
{
    const int64_t n = (int64_t)nz * ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const int x = (int)(v % nx);
        const int64_t t = v / nx;
        const int y = (int)(t % ny), z = (int)(t / ny);
        bool in = is_inlet(inl, v, z, y, x, nz, ny, nx);
        if (inl.mode == 4) in = in && cls && cls[v] < CLS_NEVER;
        parent[v + 1] = in ? 0u : (uint32_t)(v + 1);
        if (jtime) jtime[v + 1] = in ? 0 : UF_TIME_UNSET;
        // activation map: an inlet voxel is a graph node from the first radius on, whatever its class
        // (F:1265), so the link kernels need neither the inlet predicate nor the inlet mask.  Mode 4: it keeps
        // its class (a node from then on; hanging under node 0 from the start is harmless, nothing finds it earlier)
        if (acls) acls[v] = (in && inl.mode != 4) ? 0 : cls[v];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) parent[0] = 0u;
}

// Activation of one radius, in two kernels so that the pointer chasing of the unions runs with
// one thread per (voxel, neighbour) job instead of inside a sparse grid-stride scan:
//   uf_collect_kernel    : list of the voxels v in [v0, v1) with klo < cls[v] <= khi
//   uf_union_list_kernel : every listed voxel is linked to its active neighbours (active: inlet
//                          or cls <= khi).  conn: 6 (faces) or 26 (faces+edges+corners); with
//                          nz == 1 these are 4- and 8-connectivity.
// Redundant links are skipped: a y or z face pair (v, u) whose left neighbours (v-1, u-1) are
// both active is already connected through  v ~ v-1 ~ u-1 ~ u  (x links are always made, and
// the pair (v-1, u-1) is linked or skipped by the same rule: induction along the row, started
// by the first pair of every common run).  Only about one y/z link per pair of touching runs
// is left, instead of one per voxel.
__global__ void __launch_bounds__(256)
uf_collect_kernel(const uint8_t *__restrict__ cls, int klo, int khi, int64_t v0, int64_t v1,
                  uint32_t *__restrict__ list, uint32_t *__restrict__ count)
{
    // A block takes 256 x 64 consecutive voxels per round and reserves its list range with ONE
    // atomicAdd (a per-warp reservation serialises ~2M same-address atomics per radius at 1024^3).
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t block_base;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t per_block = 256 * 64;
    const int64_t nrounds = (v1 - v0 + per_block - 1) / per_block;
    const bool aligned = (((uintptr_t)(cls + v0)) & 15u) == 0;
    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
        // thread t owns the 16-voxel groups t, t + 256, t + 512, t + 768 of the round (coalesced loads)
        uint32_t mask[4];
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t v = v0 + r * per_block + ((int64_t)q * 256 + threadIdx.x) * 16;
            uint32_t m = 0;
            if (v < v1) {
                if (aligned && v + 16 <= v1) {
                    const uint4 w4 = __ldg(reinterpret_cast<const uint4 *>(cls + v));
                    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = (int)byte_of(w[i >> 2], i & 3);
                        if (c > klo && c <= khi) m |= 1u << i;
                    }
                } else {
                    for (int i = 0; i < 16 && v + i < v1; ++i) {
                        const int c = (int)cls[v + i];
                        if (c > klo && c <= khi) m |= 1u << i;
                    }
                }
            }
            mask[q] = m;
            cnt += __popc(m);
        }
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[warp] = (uint32_t)incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const uint32_t t = warp_tot[w]; warp_tot[w] = tot; tot += t; }
            block_base = tot ? atomicAdd(count, tot) : 0u;
        }
        __syncthreads();
        uint32_t base = block_base + warp_tot[warp] + (uint32_t)(incl - cnt);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t v = v0 + r * per_block + ((int64_t)q * 256 + threadIdx.x) * 16;
            uint32_t m = mask[q];
            while (m) {
                const int i = __ffs(m) - 1;
                m &= m - 1;
                list[base++] = (uint32_t)(v + i);
            }
        }
        __syncthreads();                                   // warp_tot / block_base are reused next round
    }
}

// ---- all radii at once (single-GPU loop): the voxels are bucketed by class in one pass, so that
// radius k links the slice  list[start[k] .. start[k + 1])  without scanning the class map again.
//   uf_hist_kernel   : hist[c] = number of voxels of class c (c < CLS_NEVER)
//   uf_scan_kernel   : start[] = exclusive prefix sums (257 entries), cursor[] = 0
//   uf_bucket_kernel : a block takes 256 x 64 consecutive voxels per round (a thread 64 consecutive
//                      ones), reserves its share of every class slice with one atomicAdd per class
//                      and writes runs of equal class as runs of the list (x-neighbours stay
//                      neighbours in the list, which the union kernel's loads rely on)
__global__ void __launch_bounds__(256)
uf_hist_kernel(const uint8_t *__restrict__ cls, int64_t n, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t ngroups = n / 16;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += step) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(cls) + g);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        // runs of equal bytes are counted with one atomic (all indices static: w stays in registers)
        uint32_t prev = byte_of(w[0], 0), run = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t c = byte_of(w[i >> 2], i & 3);
            if (c != prev) {
                if (prev < CLS_NEVER) atomicAdd(&h[prev], run);
                prev = c;
                run = 0;
            }
            ++run;
        }
        if (prev < CLS_NEVER) atomicAdd(&h[prev], run);
    }
    if (blockIdx.x == 0)
        for (int64_t v = ngroups * 16 + threadIdx.x; v < n; v += blockDim.x)
            if (cls[v] < CLS_NEVER) atomicAdd(&h[cls[v]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
uf_scan_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ start, uint32_t *__restrict__ cursor)
{
    __shared__ uint32_t s[256];
    s[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < 256; ++c) { const uint32_t t = s[c]; s[c] = acc; acc += t; }
        start[256] = acc;
    }
    __syncthreads();
    start[threadIdx.x] = s[threadIdx.x];
    cursor[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
uf_bucket_kernel(const uint8_t *__restrict__ cls, int64_t n, const uint32_t *__restrict__ start,
                 uint32_t *__restrict__ cursor, uint32_t *__restrict__ list)
{
    __shared__ uint32_t cnt[256];      // voxels of the round per class, then the running offset inside the reservation
    __shared__ uint32_t base[256];     // start[c] + reserved offset of this round
    const int64_t per_block = 256 * 64;
    const int64_t nrounds = (n + per_block - 1) / per_block;
    const bool aligned = (((uintptr_t)cls) & 15u) == 0;
    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
        cnt[threadIdx.x] = 0;
        __syncthreads();
        const int64_t v0 = r * per_block + (int64_t)threadIdx.x * 64;
        uint32_t w[16];
        if (aligned && v0 + 64 <= n) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 t = __ldg(reinterpret_cast<const uint4 *>(cls + v0) + q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                uint32_t t = 0;
                for (int j = 0; j < 4; ++j) {
                    const int64_t v = v0 + 4 * q + j;
                    t |= (v < n ? (uint32_t)cls[v] : CLS_BG) << (8 * j);
                }
                w[q] = t;
            }
        }
        {   // count the round per class, one atomic per run of equal bytes (static indices only)
            uint32_t prev = byte_of(w[0], 0), run = 0;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const uint32_t c = byte_of(w[i >> 2], i & 3);
                if (c != prev) {
                    if (prev < CLS_NEVER) atomicAdd(&cnt[prev], run);
                    prev = c;
                    run = 0;
                }
                ++run;
            }
            if (prev < CLS_NEVER) atomicAdd(&cnt[prev], run);
        }
        __syncthreads();
        {
            const uint32_t t = cnt[threadIdx.x];
            base[threadIdx.x] = start[threadIdx.x] + (t ? atomicAdd(&cursor[threadIdx.x], t) : 0u);
            cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        {   // write every run of equal class as a run of its list slice
            uint32_t prev = byte_of(w[0], 0), run = 0, first = 0;
#pragma unroll
            for (int i = 0; i <= 64; ++i) {
                const uint32_t c = i < 64 ? byte_of(w[(i < 64 ? i : 0) >> 2], i & 3) : 0x100u;
                if (c != prev) {
                    if (prev < CLS_NEVER) {
                        uint32_t pos = base[prev] + atomicAdd(&cnt[prev], run);
                        for (uint32_t t = 0; t < run; ++t) list[pos + t] = (uint32_t)(v0 + first + t);
                    }
                    prev = c;
                    run = 0;
                    first = (uint32_t)i;
                }
                ++run;
            }
        }
        __syncthreads();
    }
}

// list / count: the voxels to link; with `start` != NULL the slice of class khi of a bucketed list.
__global__ void __launch_bounds__(256)
uf_union_list_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int klo, int khi,
                     int conn, int nz, int ny, int nx, const uint32_t *__restrict__ list,
                     const uint32_t *__restrict__ count, uint8_t *jtime, const uint32_t *__restrict__ start)
{
    const int ndir = conn == 6 ? 6 : 26;
    if (start) {
        list += start[khi];
        count = nullptr;
    }
    const int64_t jobs = (int64_t)(start ? start[khi + 1] - start[khi] : *count) * ndir;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < jobs; j += step) {
        const uint32_t e = (uint32_t)(j / ndir);
        const int d = (int)(j - (int64_t)e * ndir);
        const uint32_t v = list[e];
        int dz, dy, dx;
        if (conn == 6) {
            const int ax = d >> 1, sg = (d & 1) ? 1 : -1;
            dz = ax == 0 ? sg : 0; dy = ax == 1 ? sg : 0; dx = ax == 2 ? sg : 0;
        } else {
            const int q = d < 13 ? d : d + 1;          // skip the centre of the 3x3x3 cube
            dz = q / 9 - 1; dy = (q / 3) % 3 - 1; dx = q % 3 - 1;
        }
        const int x = (int)(v % (uint32_t)nx);
        const uint32_t t = v / (uint32_t)nx;
        const int y = (int)(t % (uint32_t)ny), z = (int)(t / (uint32_t)ny);
        const int zz = z + dz, yy = y + dy, xx = x + dx;
        if (zz < 0 || zz >= nz || yy < 0 || yy >= ny || xx < 0 || xx >= nx) continue;
        const int64_t u = ((int64_t)zz * ny + yy) * nx + xx;
        const int cu = (int)cls[u];
        if (cu > klo && cu <= khi) {
            if (u > (int64_t)v) continue;              // both new: the pair is linked once, from the larger index
        } else if (!(cu <= khi || is_inlet(inl, u, zz, yy, xx, nz, ny, nx))) continue;
        if (dx == 0 && (dy == 0) != (dz == 0) && x > 0) {
            // y / z face pair: redundant when the pair one step to the left is active too
            const bool lv = (int)cls[v - 1u] <= khi || is_inlet(inl, (int64_t)v - 1, z, y, x - 1, nz, ny, nx);
            if (lv && ((int)cls[u - 1] <= khi || is_inlet(inl, u - 1, zz, yy, x - 1, nz, ny, nx))) continue;
        }
        int64_t w = u;
        if (dx == -1 && dy == 0 && dz == 0 && cu > klo && cu <= khi) {
            // left neighbour activated in this launch too: link to the first voxel of the run of new
            // voxels instead (every voxel of the run does), so the run becomes a star, not a chain
            // as long as the run (whose finds and the final resolve would then walk end to end)
            int xs = xx;
            for (int steps = 0; steps < 64 && xs > 0; ++steps) {
                const int c = (int)cls[w - 1];
                if (!(c > klo && c <= khi)) break;
                --w;
                --xs;
            }
        }
        uf_union(parent, v + 1u, (uint32_t)(w + 1), jtime, khi);
    }
}

// rcls[v] = k for every seed (cls <= k) that is connected to the inlets and was not marked
// at an earlier radius; *gate (monotone) is set once any voxel has ever been marked.
// A thread owns 16 consecutive voxels; the two levels parent[v], parent[parent[v]] of all its
// candidates are fetched as independent loads (after earlier compression that already decides
// almost every voxel), only deeper chains take the serial find.
__global__ void __launch_bounds__(256)
uf_mark_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, uint8_t *__restrict__ rcls,
               int k, int64_t n, int *gate)
{
    const int64_t ngroups = (n + 15) / 16;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)cls) | ((uintptr_t)rcls)) & 15u) == 0;
    int any = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += step) {
        const int64_t v = 16 * g;
        if (aligned && v + 16 <= n) {
            const uint4 cq = __ldg(reinterpret_cast<const uint4 *>(cls + v));
            uint4 rq = *reinterpret_cast<const uint4 *>(rcls + v);
            const uint32_t cw[4] = {cq.x, cq.y, cq.z, cq.w};
            uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
            uint32_t cand = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if ((int)byte_of(cw[i >> 2], i & 3) <= k && byte_of(rw[i >> 2], i & 3) == CLS_NEVER) cand |= 1u << i;
            if (cand == 0) continue;
            const volatile uint32_t *vp = parent;
            uint32_t p[16], gp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = (cand >> i & 1u) ? vp[v + i + 1] : 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) gp[i] = ((cand >> i & 1u) && p[i] != 0u) ? vp[p[i]] : 0u;   // node 0 is never read
            bool changed = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (!(cand >> i & 1u)) continue;
                uint32_t root = p[i];
                if (gp[i] != p[i]) {
                    root = uf_find(parent, gp[i]);
                    if (root != 0u) ((volatile uint32_t *)parent)[v + i + 1] = root;     // compress
                }
                if (root == 0u) {
                    rw[i >> 2] = (rw[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)k << (8 * (i & 3)));
                    changed = true;
                }
            }
            if (changed) {
                *reinterpret_cast<uint4 *>(rcls + v) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
                any = 1;
            }
        } else {
            for (int i = 0; i < 16 && v + i < n; ++i) {
                if ((int)cls[v + i] > k || rcls[v + i] != CLS_NEVER) continue;
                const uint32_t root = uf_find(parent, (uint32_t)(v + i + 1));
                if (root == 0u) { rcls[v + i] = (uint8_t)k; any = 1; }
                else ((volatile uint32_t *)parent)[v + i + 1] = root;
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, any) && lane_id() == 0 && *((volatile int *)gate) == 0) *gate = 1;
}

// One pass after the last activation (single-GPU loop):  rcls[v] = first radius index at which v
// is a seed connected to the inlets = max(cls[v], jtime[top]) where top is the child of node 0
// on v's path (see "Join times" above); CLS_NEVER for seeds whose tree never joined node 0.
// *kmin receives the smallest value written (the first radius with any reached seed).
// One thread per voxel: the lanes of a warp hold 32 consecutive voxels, which almost always sit
// in the same tree, so the walk of a warp is one chain of broadcast loads rather than 32 chains.
__global__ void __launch_bounds__(256)
uf_resolve_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, const uint8_t *__restrict__ jtime,
                  uint8_t *__restrict__ rcls, int64_t n, int *kmin)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    volatile uint32_t *vp = parent;
    uint32_t lmin = 255u;
    const bool aligned = ((((uintptr_t)cls) | ((uintptr_t)rcls)) & 3u) == 0;
    // four voxels per thread, the first three levels of their walks as independent loads (after
    // uf_compress_kernel a walk is: voxel -> chain end -> top -> node 0)
    const int64_t ngroups = aligned ? n / 4 : 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += step) {
        const uint32_t cw = __ldg(reinterpret_cast<const uint32_t *>(cls) + g);
        uint32_t out = 0;
        uint32_t x[4], p[4];
        bool live[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = byte_of(cw, j);
            live[j] = c < CLS_NEVER;
            x[j] = (uint32_t)(4 * g + j + 1);
            p[j] = live[j] ? vp[x[j]] : x[j];
        }
#pragma unroll
        for (int lvl = 0; lvl < 2; ++lvl) {
            uint32_t q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = (p[j] != 0u && p[j] != x[j]) ? vp[p[j]] : p[j];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (p[j] != 0u && p[j] != x[j]) { x[j] = p[j]; p[j] = q[j]; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = byte_of(cw, j);
            uint32_t r = c == CLS_BG ? CLS_BG : CLS_NEVER;
            if (live[j]) {
                uint32_t xx = x[j], pp = p[j];
                while (pp != 0u && pp != xx) { xx = pp; pp = vp[xx]; }          // deeper than three levels: rare
                if (pp == 0u) {
                    r = max(c, (uint32_t)jtime[xx]);
                    lmin = min(lmin, r);
                }
            }
            out |= r << (8 * j);
        }
        reinterpret_cast<uint32_t *>(rcls)[g] = out;
    }
    for (int64_t v = 4 * ngroups + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const uint32_t c = cls[v];
        uint32_t r = c == CLS_BG ? CLS_BG : CLS_NEVER;      // background stays background
        if (c < CLS_NEVER) {
            uint32_t x = (uint32_t)(v + 1), p = vp[x];
            while (p != 0u && p != x) { x = p; p = vp[x]; }
            if (p == 0u) {                                  // x: the child of node 0 on the path
                r = max(c, (uint32_t)jtime[x]);
                lmin = min(lmin, r);
            }
        }
        rcls[v] = (uint8_t)r;
    }
    lmin = __reduce_min_sync(0xFFFFFFFFu, lmin);
    if (lane_id() == 0 && lmin < 255u) atomicMin(kmin, (int)lmin);
}

// *gate = 1 once the radius index k has any reached seed (the per-radius kernels return at once
// while it is 0: F:1184 skips a radius without seeds)
__global__ void uf_gate_kernel(int *gate, const int *__restrict__ kmin, int k) { *gate = *kmin <= k ? 1 : 0; }

// ---- z-slab shards: the union-find is slab-local; connectivity through a slab face travels as
// one byte per face voxel ("this node is connected to the inlets") and is injected on the
// other side as a link to the virtual root.  Repeated by the host until no rank changes.
// out[i] = 1 if voxel i of local plane z is a graph node (inlet or cls <= k) whose root is 0.
__global__ void __launch_bounds__(256)
uf_face_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int k, int z,
               int nz, int ny, int nx, uint8_t *__restrict__ out, uint8_t *jtime = nullptr)
{
    const int64_t plane = (int64_t)ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += step) {
        const int64_t v = (int64_t)z * plane + i;
        const int y = (int)(i / nx), x = (int)(i % nx);
        const bool node = (int)cls[v] <= k || is_inlet(inl, v, z, y, x, nz, ny, nx);
        out[i] = (node && uf_find(parent, (uint32_t)(v + 1), jtime) == 0u) ? 1 : 0;
    }
}

// nb[i] != 0: the 6-neighbour of voxel i of local plane z across the slab face is connected to
// the inlets.  Every node of the plane under such a neighbour is linked to the root; *changed
// is set when that reached a component that was not connected before.
__global__ void __launch_bounds__(256)
uf_inject_kernel(uint32_t *parent, const uint8_t *__restrict__ cls, InletSpec inl, int k, int z,
                 int nz, int ny, int nx, const uint8_t *__restrict__ nb, int *changed, uint8_t *jtime = nullptr)
{
    const int64_t plane = (int64_t)ny * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += step) {
        if (nb[i] == 0) continue;
        const int64_t v = (int64_t)z * plane + i;
        const int y = (int)(i / nx), x = (int)(i % nx);
        if (!((int)cls[v] <= k || is_inlet(inl, v, z, y, x, nz, ny, nx))) continue;
        if (uf_find(parent, (uint32_t)(v + 1), jtime) == 0u) continue;
        uf_union(parent, (uint32_t)(v + 1), 0u, jtime, k);      // with join times: connected through the face at index k
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

// =====================================================================================================
// Row-rooted forest + link records (the single-GPU porosimetry loop and psb200_flood).
//
// The per-voxel job lists above spend their time enumerating (voxel, neighbour) pairs of which only a
// few per cent end in a union.  Which pairs matter can be decided from the class map alone:
//
//  * x edges.  Along a row, a voxel whose left neighbour activates no later than itself (class <=)
//    is connected to it from the moment it is active; otherwise, if its right neighbour activates
//    strictly earlier, to that one.  Those "downhill" edges form a forest of chains inside the row that
//    needs no union at all: uf_prelink_kernel writes parent[] = the chain's end (a local class minimum
//    of the row, found by pointer jumping in shared memory) together with the initialisation of the
//    forest.  A voxel is only ever used in a find after it became active, and every voxel between it
//    and its chain end is active by then, so the plain stores are exactly the unions the job kernel
//    would have made.  What remains of the x edges are the local maxima of the class along the row
//    (two chains meeting) and the cuts at the 1024-voxel segment boundaries.
//  * every other neighbour direction (dz, dy, dx): with m(x) = max(class of (z, y, x), class of
//    (z + dz, y + dy, x + dx)) -- the radius index at which the pair is active -- the pairs of one
//    direction between two rows that are active at index k form intervals of x, and all voxels of an
//    interval are x-connected inside both rows.  An interval that holds an older pair (m < k) is
//    connected through that one; so one union per interval BORN at k suffices:  the pairs with
//    m(x - 1) > m(x) <= m(x + 1), i.e. the local minima of m along x.
//
// uf_emit_kernel writes those records (and counts them per radius index and direction), uf_scan_bins_kernel turns
// the counts into slice starts, uf_sort_kernel moves the records into their slices, and radius k runs
// uf_union_rec_kernel over its slice: about 0.2 unions per voxel for all radii together instead of 6 (26) jobs
// per voxel.  (Two records per thread and iteration, to have twice the loads in flight, changed nothing: the union
// launches are bound by DRAM sector throughput on random 32-byte accesses, not by latency -- r3i.)
#define UF_SEGX 128          // a warp owns a segment: 4 consecutive voxels per lane, chains do not cross segments
#define UF_CHUNK 256         // segments per list reservation of the scatter pass
#define UF_MAXFAM 13
#define UF_NTIMES 254

// directions as (dz, dy, dx); the first 3 are 6-connectivity, all 13 the forward half of 26-connectivity
__constant__ int8_t c_uf_fam[UF_MAXFAM][3] = {{0, 0, 1},  {0, 1, 0},  {1, 0, 0},  {0, 1, -1}, {0, 1, 1},
                                              {1, 0, -1}, {1, 0, 1},  {1, 1, -1}, {1, 1, 0},  {1, 1, 1},
                                              {1, -1, -1}, {1, -1, 0}, {1, -1, 1}};

// bit j set: voxel x + j of the row is an inlet
__device__ __forceinline__ uint32_t uf_inlet_bits(const InletSpec &inl, int64_t rb, int z, int y, int x, int nz, int ny, int nx)
{
    uint32_t bits = 0;
    if (inl.mode == 2 || inl.mode == 4) {
        const uint32_t m = load4(inl.mask + rb, x, nx, 0u);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (byte_of(m, j)) bits |= 1u << j;
    } else if (inl.mode == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x + j < nx && is_inlet(inl, rb + x + j, z, y, x + j, nz, ny, nx)) bits |= 1u << j;
    }
    return bits;
}

__global__ void __launch_bounds__(256)
uf_prelink_kernel(const uint8_t *__restrict__ cls, InletSpec inl, int nz, int ny, int nx,
                  uint32_t *__restrict__ parent, uint8_t *__restrict__ jtime, uint8_t *__restrict__ acls)
{
    const uint32_t FULL = 0xFFFFFFFFu;
    const int lane = lane_id();
    const int nseg = (nx + UF_SEGX - 1) / UF_SEGX;
    const int64_t total = (int64_t)nz * ny * nseg;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    if (blockIdx.x == 0 && threadIdx.x == 0) parent[0] = 0u;
    for (int64_t seg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); seg < total; seg += nwarps) {
        const int64_t row = seg / nseg;
        const int x0 = (int)(seg - row * nseg) * UF_SEGX;
        const int z = (int)(row / ny), y = (int)(row - (int64_t)z * ny);
        const int64_t rb = row * nx;
        const int x = x0 + 4 * lane;
        uint32_t a = 0xFFFFFFFFu, inb = 0;
        if (x < nx) {
            a = load4(cls + rb, x, nx, 255u);
            inb = uf_inlet_bits(inl, rb, z, y, x, nz, ny, nx);
            if (inl.mode == 4) {
                // inlets only from their own class on: the class stays, voxels without a class are no inlets
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (byte_of(a, j) >= CLS_NEVER) inb &= ~(1u << j);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (inb >> j & 1u) a &= ~(0xFFu << (8 * j));
            }
            if (acls) store4(acls + rb, x, nx, a);
        }
        uint32_t lft = __shfl_up_sync(FULL, a, 1) >> 24, rgt = __shfl_down_sync(FULL, a, 1) & 0xFFu;
        if (lane == 0) lft = 255u;
        if (lane == 31) rgt = 255u;
        const uint32_t av[6] = {lft, byte_of(a, 0), byte_of(a, 1), byte_of(a, 2), byte_of(a, 3), rgt};
        bool pl[4], pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t ac = av[j + 1];
            const bool ok = !(inb >> j & 1u) && ac < CLS_NEVER;
            pl[j] = ok && av[j] <= ac;
            pr[j] = ok && !pl[j] && av[j + 2] < ac;
        }
        // chain ends: nearest voxel to the left that does not point left / to the right that does not point right
        int L[4], R[4], run = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) { run = pl[j] ? run : 4 * lane + j; L[j] = run; }
        int inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, off);
            if (lane >= off) inc = max(inc, t);
        }
        int exc = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) exc = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (L[j] < 0) L[j] = exc;
        run = UF_SEGX;
#pragma unroll
        for (int j = 3; j >= 0; --j) { run = pr[j] ? run : 4 * lane + j; R[j] = run; }
        inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_down_sync(FULL, inc, off);
            if (lane + off < 32) inc = min(inc, t);
        }
        exc = __shfl_down_sync(FULL, inc, 1);
        if (lane == 31) exc = UF_SEGX - 1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (R[j] >= UF_SEGX) R[j] = exc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = pl[j] ? L[j] : pr[j] ? R[j] : 4 * lane + j;
            const uint32_t rin = (__shfl_sync(FULL, inb, r >> 2) >> (r & 3)) & 1u;
            if (x + j < nx) {
                const bool hang = (inb >> j & 1u) || rin;       // inlets, and chains that end in one: children of node 0
                parent[rb + x + j + 1] = hang ? 0u : (uint32_t)(rb + x0 + r + 1);
                if (jtime) jtime[rb + x + j + 1] = hang ? 0 : UF_TIME_UNSET;
            }
        }
    }
}

// words of a row around the lane's 4 voxels: P = x-4..x-1, C = x..x+3, N = x+4..x+7 (255 outside the row).
// uf_row_load issues the loads (the lane's own word, and on lanes 0 / 31 the word beyond the segment's end),
// uf_row_finish exchanges the neighbours' words -- split so that the loads of all rows are in flight together.
__device__ __forceinline__ void uf_row_load(const uint8_t *__restrict__ rowp, bool ok, int x0, int x, int nx, int lane,
                                            uint32_t &C, uint32_t &E)
{
    C = (ok && x < nx) ? load4(rowp, x, nx, 255u) : 0xFFFFFFFFu;
    E = 0xFFFFFFFFu;
    if (lane == 0 && ok && x0 >= 4) E = load4(rowp, x0 - 4, nx, 255u);
    if (lane == 31 && ok && x + 4 < nx) E = load4(rowp, x + 4, nx, 255u);
}
__device__ __forceinline__ void uf_row_finish(uint32_t C, uint32_t E, int lane, uint32_t &P, uint32_t &N)
{
    const uint32_t FULL = 0xFFFFFFFFu;
    P = __shfl_up_sync(FULL, C, 1);
    N = __shfl_down_sync(FULL, C, 1);
    if (lane == 0) P = E;
    if (lane == 31) N = E;
}

// ---- records.  One pass over the class map (uf_emit_kernel) appends the records of a chunk of UF_CHUNK
// segments to the chunk's own region of raw[] as (voxel - first voxel of the chunk) << 13 | tag, with
// tag = index * nsub + direction (+ nfam when both voxels of the pair are new at that index: those slices come
// second, so that the unions which attach new voxels to older trees run first and the trees stay flat), and
// counts the tags; uf_scan_bins_kernel turns the counts into slice starts; uf_sort_kernel moves the records of
// every chunk into their slices of list[] as voxel ids.
#define UF_TAGBITS 13

struct UfEmit {
    uint32_t *cnt;           // shared: records per tag (all chunks of the block)
    uint32_t *fill;          // shared: records of the current chunk
    uint32_t *raw;           // the chunk's region
    uint32_t cap;
    int nfam, nsub, lane;
};

// cond: 0xFF in byte j = voxel (local id vloc + j) has a record of direction f at index byte j of m; eq: byte mask of
// the pairs whose two classes are equal
__device__ __forceinline__ void uf_emit(const UfEmit &e, uint32_t cond, uint32_t m, uint32_t eq, int f, uint32_t vloc)
{
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t c = (uint32_t)__popc(cond) >> 3;
    const uint32_t b0 = __ballot_sync(FULL, c & 1u), b1 = __ballot_sync(FULL, c & 2u), b2 = __ballot_sync(FULL, c & 4u);
    if ((b0 | b1 | b2) == 0u) return;
    const uint32_t lt = (1u << e.lane) - 1u;
    const uint32_t before = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
    const uint32_t tot = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
    uint32_t base = 0;
    if (e.lane == 0) base = atomicAdd(e.fill, tot);
    uint32_t pos = __shfl_sync(FULL, base, 0) + before;
    while (cond) {
        const int j = (__ffs(cond) - 1) >> 3;
        cond &= ~(0xFFu << (8 * j));
        uint32_t tag = byte_of(m, j) * (uint32_t)e.nsub + (uint32_t)f;
        if (e.nsub > e.nfam && byte_of(eq, j)) tag += (uint32_t)e.nfam;
        atomicAdd(&e.cnt[tag], 1u);
        if (pos < e.cap) e.raw[pos] = ((vloc + (uint32_t)j) << UF_TAGBITS) | tag;
        ++pos;
    }
}

template <int NROWS>
__global__ void __launch_bounds__(256)
uf_emit_kernel(const uint8_t *__restrict__ acls, InletSpec inl, int nz, int ny, int nx, int nfam, int nsub,
               uint32_t *__restrict__ hist, uint32_t *__restrict__ raw, uint32_t *__restrict__ chunk_count, uint32_t cap,
               int *__restrict__ overflow)
{
    extern __shared__ uint32_t uf_sm[];
    __shared__ uint32_t fill;
    const int nbins = UF_NTIMES * nsub;
    uint32_t *cnt = uf_sm;                                 // [nbins]
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int nseg = (nx + UF_SEGX - 1) / UF_SEGX;
    const int64_t total = (int64_t)nz * ny * nseg;
    const int64_t nchunks = (total + UF_CHUNK - 1) / UF_CHUNK;
    for (int i = threadIdx.x; i < nbins; i += 256) cnt[i] = 0u;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        if (threadIdx.x == 0) fill = 0u;
        __syncthreads();
        const int64_t s_end = min(total, (chunk + 1) * UF_CHUNK);
        const int64_t row0 = chunk * UF_CHUNK / nseg;
        const int64_t v0 = row0 * nx + (chunk * UF_CHUNK - row0 * nseg) * UF_SEGX;
        const UfEmit e{cnt, &fill, raw + chunk * cap, cap, nfam, nsub, lane};
        for (int64_t seg = chunk * UF_CHUNK + warp; seg < s_end; seg += 8) {
            const int64_t row = seg / nseg;
            const int x0 = (int)(seg - row * nseg) * UF_SEGX;
            const int z = (int)(row / ny), y = (int)(row - (int64_t)z * ny);
            const int64_t rb = row * nx;
            const int x = x0 + 4 * lane;
            const uint32_t vloc = (uint32_t)(rb + x - v0);
            // rows: 0 own, 1 (y+1), 2 (z+1), 3 (z+1, y+1), 4 (z+1, y-1)
            uint32_t Cw[NROWS], Ew[NROWS];
            bool okr[NROWS];
#pragma unroll
            for (int r = 0; r < NROWS; ++r) {
                const int zz = z + (r >= 2 ? 1 : 0), yy = y + (r == 1 || r == 3 ? 1 : r == 4 ? -1 : 0);
                okr[r] = zz < nz && yy >= 0 && yy < ny;
                uf_row_load(acls + ((int64_t)zz * ny + yy) * nx, okr[r], x0, x, nx, lane, Cw[r], Ew[r]);
            }
            uint32_t P, N;
            const uint32_t C = Cw[0];
            uf_row_finish(C, Ew[0], lane, P, N);
            const uint32_t Am1 = __byte_perm(P, C, 0x6543), Ap1 = __byte_perm(C, N, 0x4321);
            const uint32_t vC = __vcmpltu4(C, 0xFEFEFEFEu);                  // voxels of the lane with a class
            // ---- x edges (x, x + 1): unless one of the two points at the other in uf_prelink_kernel
            {
                uint32_t pair = vC & __vcmpltu4(Ap1, 0xFEFEFEFEu);
                uint32_t covered;
                const uint32_t sameseg = lane == 31 ? 0x00FFFFFFu : 0xFFFFFFFFu;
                const uint32_t notfirst = lane == 0 ? 0xFFFFFF00u : 0xFFFFFFFFu;
                const uint32_t le = __vcmpleu4(C, Ap1);
                const uint32_t left_ok = __vcmpleu4(Am1, C) & notfirst;
                covered = sameseg & (le | (~left_ok & ~le));
                // class 0 may be an inlet (parent = node 0, never pre-linked): the exact rule, voxel by voxel; with
                // inlets that keep their class (mode 4) any voxel may be one
                const bool any_class = inl.mode == 4;
                const uint32_t zero = any_class ? pair : (__vcmpeq4(C, 0u) | __vcmpeq4(Ap1, 0u)) & pair;
                if (zero) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (!byte_of(zero, j)) continue;
                        const uint32_t ax = byte_of(C, j), ar = byte_of(Ap1, j);
                        const int xx = x + j;
                        const bool inx = (any_class || ax == 0) && is_inlet(inl, rb + xx, z, y, xx, nz, ny, nx);
                        const bool inr = (any_class || ar == 0) && is_inlet(inl, rb + xx + 1, z, y, xx + 1, nz, ny, nx);
                        bool cov = inx && inr;
                        if (!cov && !(lane == 31 && j == 3)) {
                            const bool lok = !(lane == 0 && j == 0) && byte_of(Am1, j) <= ax;
                            cov = (!inr && ax <= ar) || (!inx && !lok && ar < ax);
                        }
                        covered = (covered & ~(0xFFu << (8 * j))) | (cov ? 0xFFu << (8 * j) : 0u);
                    }
                }
                uf_emit(e, pair & ~covered, __vmaxu4(C, Ap1), 0u, 0, vloc);
            }
            // ---- the other directions: local minima of m along x, four voxels per instruction
#pragma unroll
            for (int r = 1; r < NROWS; ++r) {
                if (!okr[r]) continue;                                       // (uniform over the warp)
                uint32_t Bp, Bn;
                const uint32_t Bc = Cw[r];
                uf_row_finish(Bc, Ew[r], lane, Bp, Bn);
                const uint32_t Bm1 = __byte_perm(Bp, Bc, 0x6543), Bp1 = __byte_perm(Bc, Bn, 0x4321);
                // the directions of this row: dx = 0 first (rows 1, 2: f = r; rows 3, 4: the middle one), then dx = -1, +1
                const int f0 = r <= 2 ? r : (r == 3 ? 8 : 11);
                const int fm = r <= 2 ? 1 + 2 * r : f0 - 1, fp = fm + (r <= 2 ? 1 : 2);
                const int ndx = NROWS > 3 ? 3 : 1;
#pragma unroll
                for (int d = 0; d < ndx; ++d) {
                    uint32_t b0, bl, br;
                    int f;
                    if (d == 0) { b0 = Bc; bl = Bm1; br = Bp1; f = f0; }
                    else if (d == 1) { b0 = Bm1; bl = __byte_perm(Bp, Bc, 0x5432); br = Bc; f = fm; }
                    else { b0 = Bp1; bl = Bc; br = __byte_perm(Bc, Bn, 0x5432); f = fp; }
                    const uint32_t m = __vmaxu4(C, b0), ml = __vmaxu4(Am1, bl), mr = __vmaxu4(Ap1, br);
                    const uint32_t cond = __vcmpgtu4(ml, m) & __vcmpgeu4(mr, m) & __vcmpltu4(m, 0xFEFEFEFEu);
                    uf_emit(e, cond, m, __vcmpeq4(C, b0), f, vloc);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            chunk_count[chunk] = min(fill, cap);
            if (fill > cap) *overflow = 1;
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += 256)
        if (cnt[b]) atomicAdd(&hist[b], cnt[b]);
}

__global__ void __launch_bounds__(256)
uf_sort_kernel(const uint32_t *__restrict__ raw, const uint32_t *__restrict__ chunk_count, uint32_t cap, int64_t nchunks,
               int nseg, int nx, int nbins, const uint32_t *__restrict__ start, uint32_t *__restrict__ cursor,
               uint32_t *__restrict__ list)
{
    extern __shared__ uint32_t uf_sm[];
    uint32_t *cnt = uf_sm, *bas = cnt + nbins;
    for (int i = threadIdx.x; i < nbins; i += 256) cnt[i] = 0u;
    __syncthreads();
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const uint32_t c = chunk_count[chunk];
        const uint32_t *rec = raw + chunk * cap;
        const int64_t row0 = chunk * UF_CHUNK / nseg;
        const uint32_t v0 = (uint32_t)(row0 * nx + (chunk * UF_CHUNK - row0 * nseg) * UF_SEGX);
        for (uint32_t i = threadIdx.x; i < c; i += 256) atomicAdd(&cnt[rec[i] & ((1u << UF_TAGBITS) - 1u)], 1u);
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += 256) {
            const uint32_t k = cnt[b];
            if (k) bas[b] = start[b] + atomicAdd(&cursor[b], k);
            cnt[b] = 0u;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c; i += 256) {
            const uint32_t r = rec[i], tag = r & ((1u << UF_TAGBITS) - 1u);
            list[bas[tag] + atomicAdd(&cnt[tag], 1u)] = v0 + (r >> UF_TAGBITS);
        }
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += 256) cnt[b] = 0u;
        __syncthreads();
    }
}

// start[] = exclusive prefix sums of hist[0 .. nbins), start[nbins] = total; cursor[] = 0
__global__ void __launch_bounds__(1024)
uf_scan_bins_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ start, uint32_t *__restrict__ cursor, int nbins)
{
    __shared__ uint32_t part[1024];
    const int per = (nbins + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    uint32_t s = 0;
    for (int b = b0; b < min(nbins, b0 + per); ++b) s += hist[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int t = 0; t < 1024; ++t) { const uint32_t c = part[t]; part[t] = acc; acc += c; }
        start[nbins] = acc;
    }
    __syncthreads();
    uint32_t acc = part[threadIdx.x];
    for (int b = b0; b < min(nbins, b0 + per); ++b) {
        start[b] = acc;
        acc += hist[b];
        cursor[b] = 0u;
    }
}

struct UfStrides { long long s[UF_MAXFAM]; };

// the unions of radius index k: list[start[k nsub] .. start[(k + 1) nsub]), direction = (slice the entry is in) mod nfam
__global__ void __launch_bounds__(256)
uf_union_rec_kernel(uint32_t *parent, const uint32_t *__restrict__ list, const uint32_t *__restrict__ start, int k,
                    int nfam, int nsub, const __grid_constant__ UfStrides st, uint8_t *jtime)
{
    __shared__ uint32_t edge[2 * UF_MAXFAM + 1];
    if (threadIdx.x <= nsub) edge[threadIdx.x] = start[k * nsub + threadIdx.x];
    __syncthreads();
    const uint32_t lo = edge[0], hi = edge[nsub];
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += step) {
        int f = 0;
        while (j >= edge[f + 1]) ++f;
        if (f >= nfam) f -= nfam;
        const uint32_t v = list[j];
        // the first two levels of both finds as independent loads: after the pre-linking almost every voxel is at
        // most one step from its tree's root (or from a child of node 0)
        const uint32_t a = v + 1u, b = (uint32_t)((long long)v + st.s[f]) + 1u;
        const volatile uint32_t *vp = parent;
        const uint32_t pa = vp[a], pb = vp[b];
        const uint32_t ga = pa ? vp[pa] : 0u, gb = pb ? vp[pb] : 0u;
        const bool da = pa == 0u || ga == 0u || ga == pa, db = pb == 0u || gb == 0u || gb == pb;     // root known
        const uint32_t ra = (pa == 0u || ga == 0u) ? 0u : pa, rb = (pb == 0u || gb == 0u) ? 0u : pb;
        if (da && db && ra == rb) continue;
        uf_union(parent, da ? ra : a, db ? rb : b, jtime, k);
    }
}

// Before the resolve pass: full path compression from every chain end.  After the unions a voxel's path is
// voxel -> (ancestors, all of them chain ends) -> top, and only the voxels that were the endpoint of a union
// had their own pointer shortened; without this pass every voxel of a chain would walk the chain end's whole
// path again.  Afterwards every chain end points at its top (the child of node 0, or the stranded root), so the
// walk of uf_resolve_kernel is at most three loads.
#define UF_CTILE 8192
__global__ void __launch_bounds__(256)
uf_compress_kernel(uint32_t *parent, const uint8_t *__restrict__ acls, int nz, int ny, int nx)
{
    // the chain ends of a tile are gathered first, so that every lane of the walks below has work
    __shared__ uint32_t queue[UF_CTILE];
    __shared__ uint32_t qn;
    const int64_t n = (int64_t)nz * ny * nx;
    const int64_t ntiles = (n + UF_CTILE - 1) / UF_CTILE;
    volatile uint32_t *vp = parent;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) qn = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < UF_CTILE; i += 256) {
            const int64_t v = tile * UF_CTILE + i;
            if (v >= n) break;
            const uint32_t ac = acls[v];
            if (ac >= CLS_NEVER) continue;
            const int x = (int)(v % nx);
            const int s = x & (UF_SEGX - 1);
            if (s > 0 && acls[v - 1] <= ac) continue;                                   // points left
            if (s + 1 < UF_SEGX && x + 1 < nx && acls[v + 1] < ac) continue;            // points right
            queue[atomicAdd(&qn, 1u)] = (uint32_t)v + 1u;
        }
        __syncthreads();
        const uint32_t cnt = qn;
        for (uint32_t q = threadIdx.x; q < cnt; q += 256) {
            const uint32_t self = queue[q];
            uint32_t xn = self, p = vp[xn];
            if (p == 0u || p == xn) continue;                                           // child of node 0 / root
            int hops = 0;
            while (true) {
                xn = p;
                p = vp[xn];
                if (p == 0u || p == xn) break;
                ++hops;
            }
            if (hops == 0) continue;
            const uint32_t top = xn;
            uint32_t yn = self;
            while (yn != top) {
                const uint32_t nxt = vp[yn];
                if (nxt == top || nxt == 0u) break;
                vp[yn] = top;
                yn = nxt;
            }
        }
        __syncthreads();
    }
}

// standalone flood after uf_compress_kernel: out[v] = 1 for the voxels of the mask (cls == 0) whose tree hangs under node 0
__global__ void __launch_bounds__(256)
uf_reach_out_kernel(const uint32_t *__restrict__ parent, const uint8_t *__restrict__ cls, uint8_t *__restrict__ out, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        uint8_t r = 0;
        if (cls[v] == 0) {
            uint32_t x = (uint32_t)(v + 1), p = parent[x];
            while (p != 0u && p != x) { x = p; p = parent[x]; }
            r = p == 0u ? 1 : 0;
        }
        out[v] = r;
    }
}
