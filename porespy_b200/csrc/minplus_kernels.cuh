// minplus_kernels.cuh -- the second / third pass of the separable transforms as a bounded
// min-plus scan:  out(u) = min_v  f(v) + (u - v)^2   along an axis that is NOT contiguous.
//
// The scan for one output starts from the bound best = f(u) and walks outwards dy = 1, 2, ...
// until dy^2 >= best: nothing farther can improve the minimum, so the work per voxel is
// O(distance to the answer) instead of O(line length), with no stack and no division (the
// lower-envelope formulation needs both).  On porous-media volumes distances are a few tens
// of voxels, so a (L + 2H)-row tile staged in shared memory serves almost every read; rows
// beyond the staged halo are fetched from global memory, so any input stays exact.
//
//   edt_minplus_kernel : y / z pass of the exact EDT (edt.edt at
//        /root/reference/src/porespy/filters/_funcs.py:1126), uint32 squared distances,
//        inner step  best = min(best, f + dy^2)  = one VIADDMNMX.U32 per voxel
//   lt_y2_kernel       : y pass of the per-radius dilation (F:1191 / F:1207), values capped at
//        T <= 32767 so two voxels share a register: one VIADDMNMX.U16x2 per two voxels
//
// Tiles are 128 columns wide (a lane owns 4 adjacent columns: conflict-free 16-/8-byte shared
// loads, 512-/128-byte coalesced global rows).
#pragma once
#include "common.cuh"

#define MP_TX 128
#define MP_WARPS 8
// Internal "infinite" squared distance: larger than any real value (3 * 32766^2), and
// MP_INF + 32766^2 still fits 32 bits, so  f + dy^2  never wraps.
#define MP_INF 0xC000FFFEu

// -------------------------------------------------------------------------- source formats
struct MpSrcU16 {                 // x-pass distances; >= 0x8000: no site in the line
    typedef uint16_t T;
    __device__ static __forceinline__ uint32_t sq(uint32_t d) { return d >= 0x8000u ? MP_INF : d * d; }
    __device__ static __forceinline__ uint4 ld4(const uint16_t *p)
    {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        return make_uint4(sq(v.x & 0xFFFFu), sq(v.x >> 16), sq(v.y & 0xFFFFu), sq(v.y >> 16));
    }
};
struct MpSrcU32 {                 // squared distances; PSB_INF: infinite
    typedef uint32_t T;
    __device__ static __forceinline__ uint32_t sq(uint32_t v) { return min(v, MP_INF); }
    __device__ static __forceinline__ uint4 ld4(const uint32_t *p)
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        return make_uint4(sq(v.x), sq(v.y), sq(v.z), sq(v.w));
    }
};

// 4 consecutive columns starting at column x of a row with `valid` columns; columns beyond the
// row read as 0 (their scan ends at once and nothing is stored for them)
template <typename Src>
__device__ __forceinline__ uint4 mp_load_row(const typename Src::T *row, int x, int64_t valid, bool vec)
{
    if (vec && x + 3 < valid) return Src::ld4(row + x);
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (x + j < valid) ? Src::sq((uint32_t)__ldg(row + x + j)) : 0u;
    return make_uint4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------ EDT pass
// Rows: index along the pass axis (n of them, `rstride` elements apart).  Columns: `nxc`
// contiguous elements.  Outer slices (blockIdx.y): `ostride` elements apart.
//   y pass: n = ny, rstride = nx,    nxc = nx,    outer = nz (ostride = ny*nx)
//   z pass: n = nz, rstride = ny*nx, nxc = ny*nx, outer = 1
// split > 0 (y pass of a z-slab shard only): rows are scattered into the send layout of the
// slab -> pencil all-to-all, so no separate pack pass is needed.
// grid.x = ceil(nxc/128) * ceil(n/L)  (row tiles fastest: neighbours share halo rows in L2).
// OUT 0: uint32 squared distance (PSB_INF when infinite); OUT 1: float32 sqrt (edt.edt's result).
template <typename Src, int OUT>
__global__ void __launch_bounds__(MP_WARPS * 32)
edt_minplus_kernel(const typename Src::T *__restrict__ src, void *__restrict__ dst, int n,
                   int64_t rstride, int64_t nxc, int64_t ostride, int L, int H, int vec,
                   uint32_t *__restrict__ gmax, int split)
{
    extern __shared__ uint4 mp_tile[];                     // [L + 2H][32]
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int nrt = (n + L - 1) / L;
    const int row0 = (int)(blockIdx.x % nrt) * L;
    const int64_t x0 = (int64_t)(blockIdx.x / nrt) * MP_TX;
    const int64_t valid = nxc - x0;                        // columns of this tile inside the row
    const typename Src::T *sbase = src + (int64_t)blockIdx.y * ostride + x0;
    const int rows = L + 2 * H;
    const int xl = 4 * lane;

    for (int r0 = warp; r0 < rows; r0 += 4 * MP_WARPS) {   // 4 independent row loads in flight
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * MP_WARPS, gr = row0 - H + r;
            v[i] = make_uint4(MP_INF, MP_INF, MP_INF, MP_INF);
            if (r < rows && gr >= 0 && gr < n) v[i] = mp_load_row<Src>(sbase + (int64_t)gr * rstride, xl, valid, vec);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * MP_WARPS;
            if (r < rows) mp_tile[r * 32 + lane] = v[i];
        }
    }
    __syncthreads();

    uint32_t lmax = 0;
    for (int ry = warp; ry < L; ry += MP_WARPS) {
        const int gr = row0 + ry;
        if (gr >= n) break;
        const int rr = ry + H;
        uint4 b = mp_tile[rr * 32 + lane];
        for (int dy = 1;; ++dy) {
            const uint32_t bm = max(max(b.x, b.y), max(b.z, b.w));
            const uint32_t d2 = (uint32_t)dy * (uint32_t)dy;
            if (d2 >= bm) break;
            const bool up_in = gr - dy >= 0, dn_in = gr + dy < n;
            if (!up_in && !dn_in) break;
            if (up_in) {
                const uint4 u = (rr - dy >= 0) ? mp_tile[(rr - dy) * 32 + lane]
                                               : mp_load_row<Src>(sbase + (int64_t)(gr - dy) * rstride, xl, valid, vec);
                b.x = __viaddmin_u32(u.x, d2, b.x); b.y = __viaddmin_u32(u.y, d2, b.y);
                b.z = __viaddmin_u32(u.z, d2, b.z); b.w = __viaddmin_u32(u.w, d2, b.w);
            }
            if (dn_in) {
                const uint4 u = (rr + dy < rows) ? mp_tile[(rr + dy) * 32 + lane]
                                                 : mp_load_row<Src>(sbase + (int64_t)(gr + dy) * rstride, xl, valid, vec);
                b.x = __viaddmin_u32(u.x, d2, b.x); b.y = __viaddmin_u32(u.y, d2, b.y);
                b.z = __viaddmin_u32(u.z, d2, b.z); b.w = __viaddmin_u32(u.w, d2, b.w);
            }
        }
        uint32_t o[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (o[j] >= MP_INF) o[j] = PSB_INF;
            if (xl + j < valid) lmax = max(lmax, o[j]);
        }
        int64_t oi = (int64_t)blockIdx.y * ostride + x0 + (int64_t)gr * rstride + xl;
        if (split > 0) {
            // y pass of a z-slab: store in all-to-all send layout [dest d][z][y - d*split][x]
            // (dest d owns rows [d*split, min(n, (d+1)*split)) of the pencil decomposition)
            const int d = gr / split, yy = gr - d * split;
            const int nyd = min(split, n - d * split);
            oi = ((int64_t)d * split * gridDim.y + (int64_t)blockIdx.y * nyd + yy) * rstride + x0 + xl;
        }
        if (OUT == 0) {
            uint32_t *orow = reinterpret_cast<uint32_t *>(dst) + oi;
            if (vec && xl + 3 < valid) *reinterpret_cast<uint4 *>(orow) = make_uint4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (xl + j < valid) orow[j] = o[j];
            }
        } else {
            float *orow = reinterpret_cast<float *>(dst) + oi;
            float f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = (o[j] == PSB_INF) ? __int_as_float(0x7F800000) : sqrtf((float)o[j]);
            if (vec && xl + 3 < valid) *reinterpret_cast<float4 *>(orow) = make_float4(f[0], f[1], f[2], f[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (xl + j < valid) orow[j] = f[j];
            }
        }
    }
    if (gmax) {
        lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
        if (lane == 0 && lmax) atomicMax(gmax, lmax);
    }
}

// --------------------------------------------------------------- per-radius y pass (uint16x2)
// gx: x-distance bytes min(d, W + 1) from xdist_kernel<XD_LT>.  reach byte
//   m = #{dz >= 0 : h + dz^2 < T} = ceil(sqrt(T - h))  where h = min_y' gx(y')^2 + (y - y')^2,
// 0 where h >= T.  Needs T <= 32767 (W <= 181).
// grid = (ceil(nx/128), ceil(ny/Ly), nz), block 256, dyn smem = (Ly + 2W) * 256 + 16 bytes.
__device__ __forceinline__ uint32_t sq_cap2(uint32_t a, uint32_t b, uint32_t W, uint32_t T)
{   // two x-distances -> packed capped squares
    const uint32_t sa = a > W ? T : a * a, sb = b > W ? T : b * b;
    return sa | (sb << 16);
}

__global__ void __launch_bounds__(256)
lt_y2_kernel(const uint8_t *__restrict__ gx, uint8_t *__restrict__ reach, int ny, int nx, uint32_t T,
             int W, int Ly, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    extern __shared__ uint4 lty2_smem[];
    uint2 *tile = reinterpret_cast<uint2 *>(lty2_smem);          // [rows][32] : 4 x u16 per lane
    const int tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const int rows = Ly + 2 * W;
    int *range = reinterpret_cast<int *>(tile + (size_t)rows * 32);   // [0] = first useful row, [1] = last
    const int x0 = blockIdx.x * MP_TX, y0 = blockIdx.y * Ly;
    const int64_t zoff = (int64_t)blockIdx.z * ny;
    if (tid == 0) { range[0] = rows; range[1] = -1; }
    __syncthreads();

    // ---- stage: thread = 16 voxels of one row (8 threads per row, 32 rows per sweep)
    const uint32_t uW = (uint32_t)W;
    int lo = rows, hi = -1;
    for (int i0 = 0; i0 < rows * 8; i0 += 4 * 256) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256 + tid, r = i >> 3, ch = i & 7;
            const int y = y0 - W + r, x = x0 + 16 * ch;
            v[u] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (r < rows && y >= 0 && y < ny && x < nx)
                v[u] = __ldg(reinterpret_cast<const uint4 *>(gx + (zoff + y) * nx + x));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256 + tid, r = i >> 3, ch = i & 7;
            if (r >= rows) continue;
            const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t s[8];
            bool useful = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t a = byte_of(w4[q], 0), b = byte_of(w4[q], 1), c = byte_of(w4[q], 2), d = byte_of(w4[q], 3);
                useful |= (a <= uW) | (b <= uW) | (c <= uW) | (d <= uW);
                s[2 * q] = sq_cap2(a, b, uW, T);
                s[2 * q + 1] = sq_cap2(c, d, uW, T);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(tile + (size_t)r * 32 + 4 * ch);
            dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
            dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
            if (useful) { lo = min(lo, r); hi = max(hi, r); }
        }
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if (lane == 0 && hi >= 0) { atomicMin(&range[0], lo); atomicMax(&range[1], hi); }
    __syncthreads();
    const int rlo = range[0], rhi = range[1];

    const uint32_t T2 = T * 0x00010001u;
    for (int ry = warp; ry < Ly; ry += 8) {
        const int y = y0 + ry;
        if (y >= ny) break;
        const int rr = ry + W;
        uint32_t outv = 0;
        // rows that can matter: within W of this row and inside [rlo, rhi]
        const int dmax = min(W, max(rr - rlo, rhi - rr));
        if (rhi >= 0 && dmax >= 0 && rr - dmax <= rhi && rr + dmax >= rlo) {
            const uint2 own = tile[rr * 32 + lane];
            uint32_t b0 = __vminu2(own.x, T2), b1 = __vminu2(own.y, T2);
            for (int dy = 1; dy <= dmax; ++dy) {
                const uint32_t m2 = __vmaxu2(b0, b1);
                const uint32_t bm = max(m2 & 0xFFFFu, m2 >> 16);
                const uint32_t d1 = (uint32_t)(dy * dy);
                if (d1 >= bm) break;
                const uint32_t d2 = d1 * 0x00010001u;
                const uint2 up = tile[(rr - dy) * 32 + lane];
                const uint2 dn = tile[(rr + dy) * 32 + lane];
                b0 = __viaddmin_u16x2(up.x, d2, b0); b1 = __viaddmin_u16x2(up.y, d2, b1);
                b0 = __viaddmin_u16x2(dn.x, d2, b0); b1 = __viaddmin_u16x2(dn.y, d2, b1);
            }
            const uint32_t h[4] = {b0 & 0xFFFFu, b0 >> 16, b1 & 0xFFFFu, b1 >> 16};
            uint32_t m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = h[j] >= T ? 0u : ceil_sqrt_small(T - h[j]);
            outv = pack4(m[0], m[1], m[2], m[3]);
        }
        const int x = x0 + 4 * lane;
        if (x < nx) *reinterpret_cast<uint32_t *>(reach + (zoff + y) * nx + x) = outv;
    }
}
