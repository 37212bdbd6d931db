// lt_kernels.cuh -- the sphere-insertion loop of porosimetry / local_thickness
// (/root/reference/src/porespy/filters/_funcs.py:1177-1209) as a bounded, all-uint8 pipeline.
//
// For one radius with integer threshold T (seeds <=> d2 >= T, fill <=> d2' < T) nothing
// farther than W = ceil(sqrt(T)) - 1 voxels from a seed can be filled, so the second EDT
// the reference computes per radius collapses to three cheap bounded steps:
//
//   classify (once per call): d2 (u32) -> class byte  k(v) = min{k : d2 >= T[k]}; seeds of
//            radius k are {class <= k} (seed sets are nested).  1 byte/voxel replaces the
//            4-byte distance map for the whole loop.
//   xy      : per (z, y-chunk, 128-column) tile: x-distance to the nearest seed from a
//            ballot-free bit mask (clz/ffs), kept in shared memory only; then the exact 2-D
//            squared distance by an outward scan over rows that stops as soon as dy^2 can
//            no longer improve the minimum; stored as the z "reach"
//            m = #{dz >= 0 : h + dz^2 < T} = ceil(sqrt(T - h))  (u8).
//   z       : fill(z) <=> exists z' : |z - z'| < m(z')  -- a max-plus cone scan, two linear
//            sweeps c = max(m, c - 1); writes the radius index where still unwritten.
//
// Background voxels (d2 == 0) can never be within sqrt(T) of a seed (a seed's open ball of
// radius^2 T is all foreground), so they are skipped everywhere.
#pragma once
#include "common.cuh"

#define LT_XT 128          // tile width in voxels (one u32 = 4 voxels per lane)
#define LT_WARPS 8
#define LT_MAX_W 253       // uint8 pipeline: T <= 254^2

// ---------------------------------------------------------------------------- classify
__device__ __forceinline__ uint32_t classify_one(uint32_t D, const uint32_t *sT, int nT)
{
    if (D == 0u) return CLS_BG;
    int lo = 0, hi = nT;               // first k with sT[k] <= D  (sT strictly descending)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sT[mid] <= D) hi = mid; else lo = mid + 1;
    }
    return lo == nT ? CLS_NEVER : (uint32_t)lo;
}

struct TArg { uint32_t t[256]; };      // thresholds, strictly descending, passed by value
struct LutArg { double v[256]; };      // radius of every index, passed by value

// Squared distances on this path are small (a few thousand on porous media), so the class comes from
// a table in shared memory (one byte load per voxel); values beyond the table take the binary search.
#define CLS_LUT 8192
__global__ void __launch_bounds__(256)
lt_classify_kernel(const uint32_t *__restrict__ d2, uint8_t *__restrict__ cls, int64_t n,
                   const __grid_constant__ TArg Targ, int nT)
{
    __shared__ uint32_t sT[256];
    __shared__ uint8_t lut[CLS_LUT];
    for (int i = threadIdx.x; i < nT; i += blockDim.x) sT[i] = Targ.t[i];
    __syncthreads();
    // d >= T[0] is class 0, so the table only spans [0, T[0])
    const uint32_t T0 = nT > 0 ? sT[0] : 0u;
    const int nlut = (int)min(T0, (uint32_t)CLS_LUT);
    for (int i = threadIdx.x; i < nlut; i += blockDim.x) lut[i] = (uint8_t)classify_one((uint32_t)i, sT, nT);
    __syncthreads();
    const int64_t n4 = n >> 2;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n4; i += step) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(d2) + i);
        const uint32_t d[4] = {v.x, v.y, v.z, v.w};
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            c[j] = d[j] >= T0 ? (nT > 0 ? 0u : classify_one(d[j], sT, nT))
                              : (d[j] < (uint32_t)nlut ? (uint32_t)lut[d[j]] : classify_one(d[j], sT, nT));
        reinterpret_cast<uint32_t *>(cls)[i] = pack4(c[0], c[1], c[2], c[3]);
    }
    for (int64_t i = (n4 << 2) + t0; i < n; i += step)
        cls[i] = (uint8_t)classify_one(d2[i], sT, nT);
}

// ------------------------------------------------------------------------------- xy pass
// grid = (ceil(nx/128), ceil(ny/Ly), nz), block = 256, dyn smem = (Ly+2W)*128 + 8*NW*4 bytes.
__global__ void __launch_bounds__(LT_WARPS * 32)
lt_xy_kernel(const uint8_t *__restrict__ cls, uint8_t *__restrict__ reach, int ny, int nx,
             int k, uint32_t T, int W, int Ly, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    extern __shared__ uint32_t lt_smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int rows = Ly + 2 * W;
    const int HW = (W + 31) >> 5, Wp = HW * 32, NW = 4 + 2 * HW;
    uint32_t *tile = lt_smem;                                // [rows][32] u32 = [rows][128] u8
    uint32_t *words = lt_smem + (size_t)rows * 32 + warp * NW;   // per-warp seed bit mask of a row
    const int x0 = blockIdx.x * LT_XT, y0 = blockIdx.y * Ly;
    const int64_t zoff = (int64_t)blockIdx.z * ny;

    // ---- phase 1: x-distance to the nearest seed for every row of the tile (+/- W halo rows)
    int any = 0;
    for (int r = warp; r < rows; r += LT_WARPS) {
        const int y = y0 - W + r;
        uint32_t packed = GX_FAR * 0x01010101u;
        if (y >= 0 && y < ny) {
            const uint8_t *row = cls + (zoff + y) * nx;
            for (int j = 0; j < NW; j += 4) {
                const int xw = x0 - Wp + 32 * j + 4 * lane;
                const uint32_t v = load4(row, xw, nx, CLS_BG);
                uint32_t nib = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) nib |= (byte_of(v, b) <= (uint32_t)k ? 1u : 0u) << b;
                uint32_t w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
                const int wi = j + (lane >> 3);
                if ((lane & 7) == 0 && wi < NW) words[wi] = w;
                any |= (w != 0u);
            }
            const uint32_t cv = load4(row, x0 + 4 * lane, nx, CLS_BG);
            __syncwarp();
            const int P0 = Wp + 4 * lane;          // window bit position of this lane's voxel 0
            const int wi = P0 >> 5;                // its 4 voxels share one mask word
            const uint32_t cur = words[wi];
            uint32_t g[4];
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4) {
                const int b = (P0 & 31) + b4;
                int dl = GX_FAR, dr = GX_FAR;
                const uint32_t ml = cur & (0xFFFFFFFFu >> (31 - b));
                if (ml) dl = b - (31 - __clz(ml));
                else
                    for (int t = 1; t <= HW; ++t) {
                        const uint32_t wv = words[wi - t];
                        if (wv) { dl = b + 32 * t - 31 + __clz(wv); break; }
                    }
                const uint32_t mr = cur & (0xFFFFFFFFu << b);
                if (mr) dr = (__ffs(mr) - 1) - b;
                else
                    for (int t = 1; t <= HW; ++t) {
                        const uint32_t wv = words[wi + t];
                        if (wv) { dr = 32 * t + (__ffs(wv) - 1) - b; break; }
                    }
                uint32_t d = (uint32_t)min(min(dl, dr), (int)GX_FAR);
                if (byte_of(cv, b4) == CLS_BG) d = GX_BG;
                g[b4] = d;
            }
            packed = pack4(g[0], g[1], g[2], g[3]);
        }
        tile[r * 32 + lane] = packed;
        __syncwarp();
    }
    const int tile_has_seed = __syncthreads_or(any);

    // ---- phase 2: exact 2-D squared distance (capped at T) for the Ly central rows -> reach
    for (int ry = warp; ry < Ly; ry += LT_WARPS) {
        const int y = y0 + ry;
        if (y >= ny) break;
        uint32_t outv = 0;
        if (tile_has_seed) {
            const int r = ry + W;
            const uint32_t v = tile[r * 32 + lane];
            uint32_t best[4];
            bool bg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a = byte_of(v, j);
                bg[j] = (a == GX_BG);
                best[j] = bg[j] ? 0u : min(T, a * a);
            }
            uint32_t bmax = max(max(best[0], best[1]), max(best[2], best[3]));
            for (int dy = 1; (uint32_t)(dy * dy) < bmax; ++dy) {
                const uint32_t up = tile[(r - dy) * 32 + lane];
                const uint32_t dn = tile[(r + dy) * 32 + lane];
                const uint32_t dy2 = (uint32_t)(dy * dy);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t cu = byte_of(up, j), cd = byte_of(dn, j);
                    best[j] = min(best[j], min(cu * cu, cd * cd) + dy2);
                }
                bmax = max(max(best[0], best[1]), max(best[2], best[3]));
            }
            uint32_t m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                m[j] = (bg[j] || best[j] >= T) ? 0u : ceil_sqrt_small(T - best[j]);
            outv = pack4(m[0], m[1], m[2], m[3]);
        }
        const int x = x0 + 4 * lane;
        if (x < nx) store4(reach + (zoff + y) * nx, x, nx, outv);
    }
}

// -------------------------------------------------------------------------------- z pass
// The volume is seen as [nz][plane] (plane = ny*nx); a thread owns VEC adjacent columns.
// grid = (ceil(plane/VEC/256), ceil(nz/LZ)), block = 256.
template <int VEC>
struct ZVec;
template <>
struct ZVec<4> {
    __device__ static __forceinline__ uint32_t ld(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
    __device__ static __forceinline__ uint32_t ldrw(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
    __device__ static __forceinline__ void st(uint8_t *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }
};
template <>
struct ZVec<1> {
    __device__ static __forceinline__ uint32_t ld(const uint8_t *p) { return __ldg(p); }
    __device__ static __forceinline__ uint32_t ldrw(const uint8_t *p) { return *p; }
    __device__ static __forceinline__ void st(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }
};

template <int LZ, int VEC>
__global__ void __launch_bounds__(256)
lt_z_kernel(const uint8_t *__restrict__ reach, const uint8_t *__restrict__ m_lo, int nlo,
            const uint8_t *__restrict__ m_hi, int nhi, uint8_t *__restrict__ idx, int nz,
            int64_t plane, int W, uint32_t val, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p >= plane) return;
    const int z0 = blockIdx.y * LZ;

    auto fetch = [&](int zz) -> uint32_t {
        if (zz < 0) return (zz >= -nlo) ? ZVec<VEC>::ld(m_lo + (int64_t)(nlo + zz) * plane + p) : 0u;
        if (zz >= nz) return (zz - nz < nhi) ? ZVec<VEC>::ld(m_hi + (int64_t)(zz - nz) * plane + p) : 0u;
        return ZVec<VEC>::ld(reach + (int64_t)zz * plane + p);
    };

    int c[VEC];
    uint32_t cf[LZ];
    // forward sweep: cones opening towards +z
#pragma unroll
    for (int j = 0; j < VEC; ++j) c[j] = 0;
    for (int zz = max(z0 - W, -nlo); zz < z0; ++zz) {
        const uint32_t v = fetch(zz);
#pragma unroll
        for (int j = 0; j < VEC; ++j) c[j] = max((int)byte_of(v, j), c[j] - 1);
    }
#pragma unroll
    for (int i = 0; i < LZ; ++i) {
        const uint32_t v = (z0 + i < nz + nhi) ? fetch(z0 + i) : 0u;
        uint32_t pk = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            c[j] = max((int)byte_of(v, j), c[j] - 1);
            pk |= (uint32_t)c[j] << (8 * j);
        }
        cf[i] = pk;
    }
    // backward sweep: cones opening towards -z, combined with the forward result
#pragma unroll
    for (int j = 0; j < VEC; ++j) c[j] = 0;
    for (int zz = min(z0 + LZ - 1 + W, nz + nhi - 1); zz >= z0 + LZ; --zz) {
        const uint32_t v = fetch(zz);
#pragma unroll
        for (int j = 0; j < VEC; ++j) c[j] = max((int)byte_of(v, j), c[j] - 1);
    }
#pragma unroll
    for (int i = LZ - 1; i >= 0; --i) {
        const int z = z0 + i;
        const uint32_t v = (z < nz + nhi) ? fetch(z) : 0u;
        uint32_t fill = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            c[j] = max((int)byte_of(v, j), c[j] - 1);
            if (c[j] > 0 || byte_of(cf[i], j) > 0) fill |= 0xFFu << (8 * j);
        }
        if (z < nz && fill) {
            uint8_t *q = idx + (int64_t)z * plane + p;
            const uint32_t old = ZVec<VEC>::ldrw(q);
            uint32_t nw = old;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (byte_of(fill, j) && byte_of(old, j) == 0) nw |= val << (8 * j);
            if (nw != old) ZVec<VEC>::st(q, nw);
        }
    }
}

// T == 1: the ball {o : |o|^2 < 1} is the single voxel, fill == seeds (F:1191 with r <= 1).
__global__ void __launch_bounds__(256)
lt_point_kernel(const uint8_t *__restrict__ cls, uint8_t *__restrict__ idx, int64_t n, int k,
                uint32_t val, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        if (cls[i] <= (uint32_t)k && idx[i] == 0) idx[i] = (uint8_t)val;
}

// ------------------------------------------------------------------------------- expand
__global__ void __launch_bounds__(256)
lt_expand_kernel(const uint8_t *__restrict__ idx, const __grid_constant__ LutArg lut_arg, int nlut,
                 double *__restrict__ out, int64_t n, int merge)
{
    __shared__ double lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (i < nlut) ? lut_arg.v[i] : 0.0;
    __syncthreads();
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const uint32_t v = idx[i];
        if (!merge) out[i] = lut[v];
        else if (v >= 1 && v < (uint32_t)nlut) out[i] = lut[v];
    }
}

__global__ void __launch_bounds__(256)
lt_mark_written_kernel(const double *__restrict__ out, uint8_t *__restrict__ idx, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        idx[i] = out[i] != 0.0 ? 255 : 0;
}

// =========================================================================================
// Streaming three-kernel form of one radius (used when nx % 16 == 0 and T <= 32767; the fused
// lt_xy_kernel above stays as the any-shape path):
//   xdist_kernel<XD_LT> (xdist_kernels.cuh) : class map -> x-distance bytes min(d, W + 1)
//   lt_y2_kernel (minplus_kernels.cuh)      : x-distance tile (+/- W halo rows) -> reach bytes
//   lt_zsweep_kernel (below)                : in-place forward cone sweep over whole z columns,
//                 then the backward sweep combined with the radius-index write
// =========================================================================================

// ------------------------------------------------------------------------------ z sweeps
// Thread = 4 adjacent columns of the [nz][plane] view, one byte lane each (SIMD-within-a-register:
// c = max(v, c - 1) is __vmaxu4(v, __vsubus4(c, 1)) for the four columns at once).  Forward sweep rewrites reach in
// place with the forward cone value (only where it differs); the backward sweep runs on those values (cones
// compose) and writes the radius index where the voxel is covered and still unwritten.
#define ZS_UNROLL 8
// 8 blocks per SM (<= 32 registers): the plane/4 threads of a 1024^2 plane then fit in ONE wave
// (262144 <= 148 * 2048); with 48 registers the launch ran 1.4 waves and its tail idled the SMs
__global__ void __launch_bounds__(256, 8)
lt_zsweep_kernel(uint8_t *__restrict__ reach, const uint8_t *__restrict__ m_lo, int nlo,
                 const uint8_t *__restrict__ m_hi, int nhi, uint8_t *__restrict__ idx, int nz,
                 int64_t plane, uint32_t val, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p >= plane) return;
    const uint32_t ONE4 = 0x01010101u, val4 = val * 0x01010101u;
    uint32_t c = 0;
    for (int z = 0; z < nlo; ++z)
        c = __vmaxu4(__ldg(reinterpret_cast<const uint32_t *>(m_lo + (int64_t)z * plane + p)), __vsubus4(c, ONE4));
    uint32_t *col = reinterpret_cast<uint32_t *>(reach + p);
    const int64_t ps = plane >> 2;                 // plane stride in u32
    int z = 0;
    for (; z + ZS_UNROLL <= nz; z += ZS_UNROLL) {
        uint32_t v[ZS_UNROLL];
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) v[i] = col[(int64_t)(z + i) * ps];
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) {
            // the forward value differs from the reach byte only inside a cone: most stores are not needed
            c = __vmaxu4(v[i], __vsubus4(c, ONE4));
            if (c != v[i]) col[(int64_t)(z + i) * ps] = c;
        }
    }
    for (; z < nz; ++z) {
        const uint32_t v = col[(int64_t)z * ps];
        c = __vmaxu4(v, __vsubus4(c, ONE4));
        if (c != v) col[(int64_t)z * ps] = c;
    }

    c = 0;
    for (int zz = nhi - 1; zz >= 0; --zz)
        c = __vmaxu4(__ldg(reinterpret_cast<const uint32_t *>(m_hi + (int64_t)zz * plane + p)), __vsubus4(c, ONE4));
    uint32_t *icol = reinterpret_cast<uint32_t *>(idx + p);
    z = nz - 1;
    for (; z - ZS_UNROLL + 1 >= 0; z -= ZS_UNROLL) {
        uint32_t v[ZS_UNROLL];
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) v[i] = col[(int64_t)(z - i) * ps];
        // cone values first (in place), then the index words of the covered planes as independent loads (see
        // lt_zsweep8_kernel: one dependent load per covered plane left the sweep waiting for DRAM round trips)
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) {
            c = __vmaxu4(v[i], __vsubus4(c, ONE4));
            v[i] = c;
        }
        uint32_t old[ZS_UNROLL];
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) old[i] = v[i] ? icol[(int64_t)(z - i) * ps] : 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < ZS_UNROLL; ++i) {
            const uint32_t m = __vcmpne4(v[i], 0u) & __vcmpeq4(old[i], 0u);      // covered and still unwritten
            if (m) icol[(int64_t)(z - i) * ps] = old[i] | (val4 & m);
        }
    }
    for (; z >= 0; --z) {
        c = __vmaxu4(col[(int64_t)z * ps], __vsubus4(c, ONE4));
        if (c) {
            const uint32_t old = icol[(int64_t)z * ps];
            const uint32_t m = __vcmpne4(c, 0u) & __vcmpeq4(old, 0u);
            if (m) icol[(int64_t)z * ps] = old | (val4 & m);
        }
    }
}

// The same sweeps with 8 columns per thread (one 8-byte load per plane) and 16 planes in flight: half the threads,
// four times the bytes in flight per thread -- the sweeps are bound by memory latency and bandwidth, not by the ALU.
#define ZS2_UNROLL 16
#define ZS2_IDXB 8           // planes per batch of the backward sweep (reach words + index words: register budget 72)
// 128-thread blocks, 7 per SM (<= 73 registers): the plane/8 threads of a 1024^2 plane are 1024 blocks <= 148 * 7
__global__ void __launch_bounds__(128, 7)
lt_zsweep8_kernel(uint8_t *__restrict__ reach, const uint8_t *__restrict__ m_lo, int nlo,
                  const uint8_t *__restrict__ m_hi, int nhi, uint8_t *__restrict__ idx, int nz,
                  int64_t plane, uint32_t val, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (p >= plane) return;
    const uint32_t ONE4 = 0x01010101u, val4 = val * 0x01010101u;
    uint32_t c0 = 0, c1 = 0;
    for (int z = 0; z < nlo; ++z) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(m_lo + (int64_t)z * plane + p));
        c0 = __vmaxu4(v.x, __vsubus4(c0, ONE4));
        c1 = __vmaxu4(v.y, __vsubus4(c1, ONE4));
    }
    uint2 *col = reinterpret_cast<uint2 *>(reach + p);
    const int64_t ps = plane >> 3;                 // plane stride in uint2
    int z = 0;
    for (; z + ZS2_UNROLL <= nz; z += ZS2_UNROLL) {
        uint2 v[ZS2_UNROLL];
#pragma unroll
        for (int i = 0; i < ZS2_UNROLL; ++i) v[i] = col[(int64_t)(z + i) * ps];
#pragma unroll
        for (int i = 0; i < ZS2_UNROLL; ++i) {
            c0 = __vmaxu4(v[i].x, __vsubus4(c0, ONE4));
            c1 = __vmaxu4(v[i].y, __vsubus4(c1, ONE4));
            if (c0 != v[i].x || c1 != v[i].y) col[(int64_t)(z + i) * ps] = make_uint2(c0, c1);
        }
    }
    for (; z < nz; ++z) {
        const uint2 v = col[(int64_t)z * ps];
        c0 = __vmaxu4(v.x, __vsubus4(c0, ONE4));
        c1 = __vmaxu4(v.y, __vsubus4(c1, ONE4));
        if (c0 != v.x || c1 != v.y) col[(int64_t)z * ps] = make_uint2(c0, c1);
    }
    c0 = c1 = 0;
    for (int zz = nhi - 1; zz >= 0; --zz) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(m_hi + (int64_t)zz * plane + p));
        c0 = __vmaxu4(v.x, __vsubus4(c0, ONE4));
        c1 = __vmaxu4(v.y, __vsubus4(c1, ONE4));
    }
    uint2 *icol = reinterpret_cast<uint2 *>(idx + p);
    auto commit = [&](int64_t zrow) {
        if (c0 | c1) {
            const uint2 old = icol[zrow * ps];
            const uint32_t m0 = __vcmpne4(c0, 0u) & __vcmpeq4(old.x, 0u), m1 = __vcmpne4(c1, 0u) & __vcmpeq4(old.y, 0u);
            if (m0 | m1) icol[zrow * ps] = make_uint2(old.x | (val4 & m0), old.y | (val4 & m1));
        }
    };
    z = nz - 1;
    for (; z - ZS2_IDXB + 1 >= 0; z -= ZS2_IDXB) {
        // batches of ZS2_IDXB planes: the reach words, then the cone values (ALU only, in place), then the index words
        // of the covered planes as independent loads -- a load per plane inside the dependent chain left the sweep
        // waiting for one DRAM round trip per covered plane
        uint2 v[ZS2_IDXB], old[ZS2_IDXB];
#pragma unroll
        for (int i = 0; i < ZS2_IDXB; ++i) v[i] = col[(int64_t)(z - i) * ps];
#pragma unroll
        for (int i = 0; i < ZS2_IDXB; ++i) {
            c0 = __vmaxu4(v[i].x, __vsubus4(c0, ONE4));
            c1 = __vmaxu4(v[i].y, __vsubus4(c1, ONE4));
            v[i] = make_uint2(c0, c1);
        }
#pragma unroll
        for (int i = 0; i < ZS2_IDXB; ++i)
            old[i] = (v[i].x | v[i].y) ? icol[(int64_t)(z - i) * ps] : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
        for (int i = 0; i < ZS2_IDXB; ++i) {
            const uint32_t m0 = __vcmpne4(v[i].x, 0u) & __vcmpeq4(old[i].x, 0u);
            const uint32_t m1 = __vcmpne4(v[i].y, 0u) & __vcmpeq4(old[i].y, 0u);
            if (m0 | m1) icol[(int64_t)(z - i) * ps] = make_uint2(old[i].x | (val4 & m0), old[i].y | (val4 & m1));
        }
    }
    for (; z >= 0; --z) {
        const uint2 v = col[(int64_t)z * ps];
        c0 = __vmaxu4(v.x, __vsubus4(c0, ONE4));
        c1 = __vmaxu4(v.y, __vsubus4(c1, ONE4));
        commit(z);
    }
}

// ---- z-slab shards: what a neighbour needs from this slab's reach bytes for ITS cone sweeps is one number per
// column -- the cone value arriving at the shared face, c = max_j (m_j - j) over the `depth` planes next to the
// face (j = 0: the face plane).  The neighbour feeds it to its sweep as a single halo plane (nlo / nhi = 1):
// W planes of halo traffic per radius become one.
// side 0: cone running down through planes depth-1 .. 0 (for the lower neighbour); side 1: up through
// nz-depth .. nz-1 (for the upper neighbour).
__global__ void __launch_bounds__(256)
lt_halo_cone_kernel(const uint8_t *__restrict__ reach, int nz, int64_t plane, int depth, int side,
                    uint8_t *__restrict__ out)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (plane & 3) == 0 && ((((uintptr_t)reach) | ((uintptr_t)out)) & 3u) == 0;
    if (vec) {
        for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; p < plane; p += step * 4) {
            uint32_t c = 0;
            for (int j = depth - 1; j >= 0; --j) {
                const int z = side ? nz - 1 - j : j;
                c = __vmaxu4(__ldg(reinterpret_cast<const uint32_t *>(reach + (int64_t)z * plane + p)), __vsubus4(c, 0x01010101u));
            }
            *reinterpret_cast<uint32_t *>(out + p) = c;
        }
    } else {
        for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += step) {
            uint32_t c = 0;
            for (int j = depth - 1; j >= 0; --j) {
                const int z = side ? nz - 1 - j : j;
                const uint32_t v = reach[(int64_t)z * plane + p];
                c = max(v, c ? c - 1u : 0u);
            }
            out[p] = (uint8_t)c;
        }
    }
}
