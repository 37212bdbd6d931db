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
#define MP_TS 34                // shared-memory tile row stride of the 16-bit kernels in uint2 (272 bytes, see lt_y2_kernel)
#define MP_WARPS 8
// Internal "infinite" squared distance: larger than any real value (3 * 32766^2), and
// MP_INF + 32766^2 still fits 32 bits, so  f + dy^2  never wraps.
#define MP_INF 0xC000FFFEu

// -------------------------------------------------------------------------- source formats
struct MpSrcU16 {                 // x-pass distances; >= 0x8000: no site in the line
    typedef uint16_t T;
    typedef uint2 Raw;            // 4 columns as loaded
    __device__ static __forceinline__ uint32_t sq(uint32_t d) { return d >= 0x8000u ? MP_INF : d * d; }
    __device__ static __forceinline__ Raw ldraw(const uint16_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    __device__ static __forceinline__ uint4 cvt(const Raw &v)
    {
        return make_uint4(sq(v.x & 0xFFFFu), sq(v.x >> 16), sq(v.y & 0xFFFFu), sq(v.y >> 16));
    }
    __device__ static __forceinline__ uint4 ld4(const uint16_t *p) { return cvt(ldraw(p)); }
};
struct MpSrcU32 {                 // squared distances; PSB_INF: infinite
    typedef uint32_t T;
    typedef uint4 Raw;
    __device__ static __forceinline__ uint32_t sq(uint32_t v) { return min(v, MP_INF); }
    __device__ static __forceinline__ Raw ldraw(const uint32_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    __device__ static __forceinline__ uint4 cvt(const Raw &v) { return make_uint4(sq(v.x), sq(v.y), sq(v.z), sq(v.w)); }
    __device__ static __forceinline__ uint4 ld4(const uint32_t *p) { return cvt(ldraw(p)); }
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
// gmax (optional) receives the maximum over the rows [mrow0, mrow1) of the pass axis only (a z-slab shard
// computes on its slab plus halo planes, but the maximum that sets the radii is the one of its own planes).
#define MP_R 4                 // rows per lane: a lane owns a 4 (rows) x 4 (columns) block of outputs

__device__ __forceinline__ void mp_relax(uint4 &b, const uint4 &u, uint32_t d2)
{
    b.x = __viaddmin_u32(u.x, d2, b.x); b.y = __viaddmin_u32(u.y, d2, b.y);
    b.z = __viaddmin_u32(u.z, d2, b.z); b.w = __viaddmin_u32(u.w, d2, b.w);
}
__device__ __forceinline__ uint32_t mp_max4(const uint4 &b) { return max(max(b.x, b.y), max(b.z, b.w)); }

// Register tiling along the pass axis: the two rows fetched at step dy (one above, one below the
// 4-row block) serve all 4 output rows with the offsets (dy + i)^2 / (dy + 3 - i)^2, so the
// shared-memory traffic and the loop overhead per output are a quarter of the one-row form.
template <typename Src, int OUT>
__global__ void __launch_bounds__(MP_WARPS * 32)
edt_minplus_kernel(const typename Src::T *__restrict__ src, void *__restrict__ dst, int n,
                   int64_t rstride, int64_t nxc, int64_t ostride, int L, int H, int vec,
                   uint32_t *__restrict__ gmax, int split, const int *__restrict__ gate,
                   int64_t tiles_x, int nouter, int mrow0, int mrow1)
{
    if (gate && *gate == 0) return;        // the 16-bit form of the pass (below) resolved every voxel
    extern __shared__ uint4 mp_tile[];                     // [roundup4(L) + 2H][32]
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int nrt = (n + L - 1) / L;
    uint32_t lmax = 0;
    // blocks walk the (tiles_x, nouter) tile grid: the launch is one wave of blocks, so a launch that
    // is gated off costs microseconds
    for (int64_t tile_id = blockIdx.x; tile_id < tiles_x * nouter; tile_id += gridDim.x) {
    const int64_t bx = tile_id % tiles_x;
    const int by = (int)(tile_id / tiles_x);
    const int row0 = (int)(bx % nrt) * L;
    const int64_t x0 = (int64_t)(bx / nrt) * MP_TX;
    const int64_t valid = nxc - x0;                        // columns of this tile inside the row
    const typename Src::T *sbase = src + (int64_t)by * ostride + x0;
    const int Lr = (L + MP_R - 1) & ~(MP_R - 1);
    const int rows = Lr + 2 * H;
    const uint4 INF4 = make_uint4(MP_INF, MP_INF, MP_INF, MP_INF);

    for (int r0 = warp; r0 < rows; r0 += 4 * MP_WARPS) {   // 4 independent row loads in flight
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * MP_WARPS, gr = row0 - H + r;
            v[i] = INF4;
            if (r < rows && gr >= 0 && gr < n) v[i] = mp_load_row<Src>(sbase + (int64_t)gr * rstride, 4 * lane, valid, vec);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * MP_WARPS;
            if (r < rows) mp_tile[r * 32 + lane] = v[i];
        }
    }
    __syncthreads();

    // Warp footprint: 32 columns x 16 rows (lane = 8 column groups x 4 row blocks) -- compact, so
    // the lanes of a warp see similar distances; every quarter-warp still reads 128 contiguous
    // bytes of one tile row (conflict-free LDS.128).  Block: 4 warps across x, 2 down.
    const int cg = (warp & 3) * 8 + (lane & 7);            // column group (uint4) inside the tile row
    const int xl = 4 * cg;
    const uint4 *col = mp_tile + cg;
    for (int ry = ((warp >> 2) * 4 + (lane >> 3)) * MP_R; ry < L; ry += 2 * 4 * MP_R) {
        const int gr = row0 + ry;                          // first output row of the block
        if (gr >= n) break;                                // (per lane: no warp-wide operation inside)
        const int rr = ry + H;
        uint4 O[MP_R], B[MP_R];
#pragma unroll
        for (int i = 0; i < MP_R; ++i) O[i] = col[(rr + i) * 32];
#pragma unroll
        for (int i = 0; i < MP_R; ++i) {
            B[i] = O[i];
#pragma unroll
            for (int j = 0; j < MP_R; ++j)
                if (j != i) mp_relax(B[i], O[j], (uint32_t)((i - j) * (i - j)));
            if (gr + i >= n) B[i] = make_uint4(0u, 0u, 0u, 0u);      // rows past the end: nothing to do
        }
        // fast loop: both fetched rows are inside the staged tile (rows outside the volume hold INF)
        const int dlim = min(rr, rows - 1 - (rr + MP_R - 1));
        int dy = 1;
        bool done = false;
        // two steps per termination test (a step past the bound only relaxes with valid candidates)
        while (dy <= dlim) {
            const uint32_t bm = max(max(mp_max4(B[0]), mp_max4(B[1])), max(mp_max4(B[2]), mp_max4(B[3])));
            if ((uint32_t)(dy * dy) >= bm) { done = true; break; }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                if (dy > dlim) break;
                const uint4 top = col[(rr - dy) * 32], bot = col[(rr + MP_R - 1 + dy) * 32];
                uint32_t of[MP_R];
#pragma unroll
                for (int i = 0; i < MP_R; ++i) of[i] = (uint32_t)((dy + i) * (dy + i));
#pragma unroll
                for (int i = 0; i < MP_R; ++i) {
                    mp_relax(B[i], top, of[i]);
                    mp_relax(B[i], bot, of[MP_R - 1 - i]);
                }
                ++dy;
            }
        }
        // slow loop (rare): rows beyond the staged halo come from global memory
        while (!done) {
            const uint32_t bm = max(max(mp_max4(B[0]), mp_max4(B[1])), max(mp_max4(B[2]), mp_max4(B[3])));
            const bool up_in = gr - dy >= 0, dn_in = gr + MP_R - 1 + dy < n;
            if ((uint32_t)dy * (uint32_t)dy >= bm || (!up_in && !dn_in)) break;
            if (up_in) {
                const uint4 top = (rr - dy >= 0) ? col[(rr - dy) * 32]
                                                 : mp_load_row<Src>(sbase + (int64_t)(gr - dy) * rstride, xl, valid, vec);
#pragma unroll
                for (int i = 0; i < MP_R; ++i) mp_relax(B[i], top, (uint32_t)(dy + i) * (uint32_t)(dy + i));
            }
            if (dn_in) {
                const int rb = rr + MP_R - 1 + dy;
                const uint4 bot = (rb < rows) ? col[rb * 32]
                                              : mp_load_row<Src>(sbase + (int64_t)(gr + MP_R - 1 + dy) * rstride, xl, valid, vec);
#pragma unroll
                for (int i = 0; i < MP_R; ++i)
                    mp_relax(B[i], bot, (uint32_t)(dy + MP_R - 1 - i) * (uint32_t)(dy + MP_R - 1 - i));
            }
            ++dy;
        }
        // ---- store the block
#pragma unroll
        for (int i = 0; i < MP_R; ++i) {
            const int g = gr + i;
            if (g >= n || ry + i >= L) continue;
            int64_t oi = (int64_t)by * ostride + x0 + (int64_t)g * rstride + xl;
            if (split > 0) {
                // y pass of a z-slab: all-to-all send layout [dest d][z][y - d*split][x]
                // (dest d owns rows [d*split, min(n, (d+1)*split)) of the pencil decomposition)
                const int d = g / split, yy = g - d * split;
                const int nyd = min(split, n - d * split);
                oi = ((int64_t)d * split * nouter + (int64_t)by * nyd + yy) * rstride + x0 + xl;
            }
            const uint32_t o[4] = {B[i].x, B[i].y, B[i].z, B[i].w};
            if (g >= mrow0 && g < mrow1) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (xl + j < valid) lmax = max(lmax, o[j]);
            }
            if (OUT == 0) {
                // infinite values stay >= MP_INF here; edt_fix_inf_kernel maps them to PSB_INF
                // afterwards, and only when the running max says there are any
                uint32_t *orow = reinterpret_cast<uint32_t *>(dst) + oi;
                if (vec && xl + 3 < valid) *reinterpret_cast<uint4 *>(orow) = B[i];
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (xl + j < valid) orow[j] = o[j];
                }
            } else {
                float *orow = reinterpret_cast<float *>(dst) + oi;
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = (o[j] >= MP_INF) ? __int_as_float(0x7F800000) : sqrtf((float)o[j]);
                if (vec && xl + 3 < valid) *reinterpret_cast<float4 *>(orow) = make_float4(f[0], f[1], f[2], f[3]);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (xl + j < valid) orow[j] = f[j];
                }
            }
        }
    }
    __syncthreads();                                       // the tile is restaged by the next round
    }
    if (gmax) {
        lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
        if (lane == 0 && lmax) atomicMax(gmax, min(lmax, MP_INF));
    }
}

// ------------------------------------------------- EDT pass, two voxels per instruction
// The same bounded scan with 16-bit lanes (VIADDMNMX.U16x2: half the issue slots and half the
// shared-memory bytes per voxel).  Values and offsets are capped at MP16_CAP = 32767, so a sum
// never wraps, and a result below the cap is exact: its minimising candidate has value and
// offset below the cap (represented exactly), and every other candidate's capped sum is
// >= min(its true sum, cap).  A result >= cap (a voxel 181+ voxels away from the background
// along this and the previous axes -- never on porous media) raises *overflow, and the
// uint32 kernel above, launched right behind on the same stream and gated on that flag,
// recomputes the pass.
#define MP16_CAP 0x7FFFu
#define MP16_WARPS 8
struct MpTrue { static constexpr bool value = true; };
struct MpFalse { static constexpr bool value = false; };

__device__ __forceinline__ uint2 mp16_pack(const uint4 &v)
{
    return make_uint2(min(v.x, MP16_CAP) | (min(v.y, MP16_CAP) << 16),
                      min(v.z, MP16_CAP) | (min(v.w, MP16_CAP) << 16));
}
__device__ __forceinline__ uint32_t mp16_off(int d) { return min((uint32_t)(d * d), MP16_CAP) * 0x00010001u; }

// grid / tile geometry as edt_minplus_kernel; dyn smem = rows * MP_TS * 8 + (H + 2) * 16 bytes.
// FOOT: warp footprint of the scan, 0 = 64 columns x 8 rows, 1 = 32 x 16 (see lt_y2_kernel).
template <typename Src, int OUT, int FOOT>
__global__ void __launch_bounds__(MP16_WARPS * 32, 4)
edt_minplus16_kernel(const typename Src::T *__restrict__ src, void *__restrict__ dst, int n,
                     int64_t rstride, int64_t nxc, int64_t ostride, int L, int H, int vec,
                     uint32_t *__restrict__ gmax, int split, int *__restrict__ overflow, int mrow0, int mrow1)
{
    extern __shared__ uint4 mp16_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const int nrt = (n + L - 1) / L;
    const int row0 = (int)(blockIdx.x % nrt) * L;
    const int64_t x0 = (int64_t)(blockIdx.x / nrt) * MP_TX;
    const int64_t valid = nxc - x0;
    const typename Src::T *sbase = src + (int64_t)blockIdx.y * ostride + x0;
    const int Lr = (L + 3) & ~3;
    const int rows = Lr + 2 * H;
    uint2 *tile = reinterpret_cast<uint2 *>(mp16_smem);                   // [rows][MP_TS]: 4 x u16 per entry
    uint4 *offt = reinterpret_cast<uint4 *>(tile + (size_t)rows * MP_TS);    // [H + 2]: packed capped (d + i)^2, i = 0..3
    for (int d = tid; d < H + 2; d += MP16_WARPS * 32)
        offt[d] = make_uint4(mp16_off(d), mp16_off(d + 1), mp16_off(d + 2), mp16_off(d + 3));
    const uint4 INF4 = make_uint4(MP_INF, MP_INF, MP_INF, MP_INF);
    // staging is latency-bound (the whole block waits for it): 8 independent row loads in flight per warp
    if (vec && 4 * lane + 3 < valid) {
        for (int r0 = warp; r0 < rows; r0 += 8 * MP16_WARPS) {
            typename Src::Raw raw[8];
            bool in[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + i * MP16_WARPS, gr = row0 - H + r;
                in[i] = r < rows && gr >= 0 && gr < n;
                if (in[i]) raw[i] = Src::ldraw(sbase + (int64_t)gr * rstride + 4 * lane);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + i * MP16_WARPS;
                if (r < rows) tile[r * MP_TS + lane] = mp16_pack(in[i] ? Src::cvt(raw[i]) : INF4);
            }
        }
    } else {
        for (int r0 = warp; r0 < rows; r0 += MP16_WARPS) {
            const int gr = row0 - H + r0;
            uint4 v = INF4;
            if (gr >= 0 && gr < n) v = mp_load_row<Src>(sbase + (int64_t)gr * rstride, 4 * lane, valid, vec);
            tile[r0 * MP_TS + lane] = mp16_pack(v);
        }
    }
    __syncthreads();

    // A lane owns 4 rows x 4 columns; warp footprint 64 columns x 8 rows (16 column groups x 2 row
    // blocks: every half-warp reads 128 contiguous bytes of one tile row) or 32 x 16 (FOOT 1).
    // FULL (decided per block): the tile lies inside the volume, rows and columns, and stores go to the plain layout,
    // so the per-row / per-column bounds tests and the 64-bit index arithmetic of the general form drop out (they were
    // 45 % of the executed instructions: profiles/r2g_sass_hotspots.txt)
    uint32_t lmax = 0, amax = 0;
    const int cq = FOOT ? (warp & 3) * 8 + (lane & 7) : (warp & 1) * 16 + (lane & 15);
    const int xl = 4 * cq;
    const int ry0 = FOOT ? ((warp >> 2) * 4 + (lane >> 3)) * 4 : ((warp >> 1) * 2 + (lane >> 4)) * 4;
    auto body = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        for (int ry = ry0; ry < L; ry += 32) {
            const int gr = row0 + ry;
            if (!FULL && gr >= n) break;
            const int rr = ry + H;
            uint2 O[4];
            uint32_t B0[4], B1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) O[i] = tile[(rr + i) * MP_TS + cq];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                B0[i] = O[i].x;
                B1[i] = O[i].y;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j != i) {
                        const uint32_t d = (uint32_t)((i - j) * (i - j)) * 0x00010001u;
                        B0[i] = __viaddmin_u16x2(O[j].x, d, B0[i]);
                        B1[i] = __viaddmin_u16x2(O[j].y, d, B1[i]);
                    }
                if (!FULL && gr + i >= n) B0[i] = B1[i] = 0u;                // rows past the end: nothing to do
            }
            const int dlim = FULL ? H : min(min(rr, rows - 1 - (rr + 3)), H);
            int dy = 1;
            bool done = false;
            // fast loop: both fetched rows are inside the staged tile (rows outside the volume hold the cap);
            // two steps per termination test (a step past the bound only relaxes with valid candidates)
            while (dy <= dlim) {
                const uint32_t m2 = __vmaxu2(__vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1])),
                                             __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3])));
                const uint32_t bm = max(m2 & 0xFFFFu, m2 >> 16);
                if ((uint32_t)(dy * dy) >= bm) { done = true; break; }
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    if (dy > dlim) break;
                    const uint2 top = tile[(rr - dy) * MP_TS + cq];
                    const uint2 bot = tile[(rr + 3 + dy) * MP_TS + cq];
                    const uint4 o4 = offt[dy];
                    const uint32_t of[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        B0[i] = __viaddmin_u16x2(top.x, of[i], B0[i]); B1[i] = __viaddmin_u16x2(top.y, of[i], B1[i]);
                        B0[i] = __viaddmin_u16x2(bot.x, of[3 - i], B0[i]); B1[i] = __viaddmin_u16x2(bot.y, of[3 - i], B1[i]);
                    }
                    ++dy;
                }
            }
            // slow loop (rare): rows beyond the staged halo come from global memory
            while (!done) {
                const uint32_t m2 = __vmaxu2(__vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1])),
                                             __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3])));
                const uint32_t bm = max(m2 & 0xFFFFu, m2 >> 16);
                const bool up_in = gr - dy >= 0, dn_in = gr + 3 + dy < n;
                if ((uint32_t)dy * (uint32_t)dy >= bm || (!up_in && !dn_in)) break;
                const uint32_t of[4] = {mp16_off(dy), mp16_off(dy + 1), mp16_off(dy + 2), mp16_off(dy + 3)};
                if (up_in) {
                    const uint2 top = (rr - dy >= 0) ? tile[(rr - dy) * MP_TS + cq]
                                                     : mp16_pack(mp_load_row<Src>(sbase + (int64_t)(gr - dy) * rstride, xl, valid, vec));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        B0[i] = __viaddmin_u16x2(top.x, of[i], B0[i]); B1[i] = __viaddmin_u16x2(top.y, of[i], B1[i]);
                    }
                }
                if (dn_in) {
                    const int rb = rr + 3 + dy;
                    const uint2 bot = (rb < rows) ? tile[rb * MP_TS + cq]
                                                  : mp16_pack(mp_load_row<Src>(sbase + (int64_t)(gr + 3 + dy) * rstride, xl, valid, vec));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        B0[i] = __viaddmin_u16x2(bot.x, of[3 - i], B0[i]); B1[i] = __viaddmin_u16x2(bot.y, of[3 - i], B1[i]);
                    }
                }
                ++dy;
            }
            // ---- store the block
            if (FULL) {
                char *orow = reinterpret_cast<char *>(dst) + 4 * ((int64_t)blockIdx.y * ostride + x0 + (int64_t)gr * rstride + xl);
                const uint32_t r01 = __vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1]));
                const uint32_t r23 = __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3]));
                const uint32_t m2 = __vmaxu2(r01, r23);
                amax = max(amax, max(m2 & 0xFFFFu, m2 >> 16));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t o[4] = {B0[i] & 0xFFFFu, B0[i] >> 16, B1[i] & 0xFFFFu, B1[i] >> 16};
                    if (gr + i >= mrow0 && gr + i < mrow1) {
                        const uint32_t mi = __vmaxu2(B0[i], B1[i]);
                        lmax = max(lmax, max(mi & 0xFFFFu, mi >> 16));
                    }
                    if (OUT == 0) *reinterpret_cast<uint4 *>(orow) = make_uint4(o[0], o[1], o[2], o[3]);
                    else *reinterpret_cast<float4 *>(orow) = make_float4(sqrtf((float)o[0]), sqrtf((float)o[1]), sqrtf((float)o[2]), sqrtf((float)o[3]));
                    orow += 4 * rstride;
                }
                continue;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int g = gr + i;
                if (g >= n || ry + i >= L) continue;
                int64_t oi = (int64_t)blockIdx.y * ostride + x0 + (int64_t)g * rstride + xl;
                if (split > 0) {
                    // y pass of a z-slab: all-to-all send layout [dest d][z][y - d*split][x]
                    const int d = g / split, yy = g - d * split;
                    const int nyd = min(split, n - d * split);
                    oi = ((int64_t)d * split * gridDim.y + (int64_t)blockIdx.y * nyd + yy) * rstride + x0 + xl;
                }
                const uint32_t o[4] = {B0[i] & 0xFFFFu, B0[i] >> 16, B1[i] & 0xFFFFu, B1[i] >> 16};
                {   // row maximum: every row feeds the overflow test, the rows [mrow0, mrow1) feed gmax
                    uint32_t rmax = 0;
                    if (xl + 3 < valid) {
                        const uint32_t m2 = __vmaxu2(B0[i], B1[i]);
                        rmax = max(m2 & 0xFFFFu, m2 >> 16);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (xl + j < valid) rmax = max(rmax, o[j]);
                    }
                    amax = max(amax, rmax);
                    if (g >= mrow0 && g < mrow1) lmax = max(lmax, rmax);
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
                    for (int j = 0; j < 4; ++j) f[j] = sqrtf((float)o[j]);
                    if (vec && xl + 3 < valid) *reinterpret_cast<float4 *>(orow) = make_float4(f[0], f[1], f[2], f[3]);
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (xl + j < valid) orow[j] = f[j];
                    }
                }
            }
        }
    };
    const bool full = vec && valid >= MP_TX && row0 + L <= n && (L & 3) == 0 && split == 0;
    if (full) body(MpTrue());
    else body(MpFalse());
    if (__any_sync(0xFFFFFFFFu, amax >= MP16_CAP) && lane == 0) *overflow = 1;
    if (gmax) {
        lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
        if (lane == 0 && lmax) atomicMax(gmax, lmax);
    }
}

// values >= MP_INF (infinite: no background in the volume / plane) -> PSB_INF
// Runs only when the running max of the last pass says there are any (device-side gate, no
// host round trip); also turns the max itself into PSB_INF.
__global__ void __launch_bounds__(256)
edt_fix_inf_kernel(uint32_t *__restrict__ d2, int64_t n, uint32_t *__restrict__ gmax)
{
    if (*reinterpret_cast<volatile uint32_t *>(gmax) < MP_INF) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        if (d2 && d2[i] >= MP_INF) d2[i] = PSB_INF;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(gmax, PSB_INF);
}

// --------------------------------------------------------------- per-radius y pass (uint16x2)
// gx: x-distance bytes min(d, W + 1) from xdist_kernel<XD_LT>.  reach byte
//   m = #{dz >= 0 : h + dz^2 < T} = ceil(sqrt(T - h))  where h = min_y' gx(y')^2 + (y - y')^2,
// 0 where h >= T.  Needs T <= 32767 (W <= 181).
// Measured and NOT adopted (r2b): letting background voxels (never filled: a seed's open ball holds no
// background voxel) start with best = 0 so that they do not prolong the scan of their 4 x 4 block.  It
// removes 40 % of the per-BLOCK scan steps, but a warp scans until the last of its 32 blocks is done, and
// nearly every 64 x 8 patch holds an unreachable pore voxel: lt_y 2.49 -> 2.79 ms at T = 344 (the marking
// of background voxels in the x pass cost another 0.08 ms per radius).
// grid = (ceil(nx/128), ceil(ny/Ly), nz), block 256, dyn smem: see lt_y2_smem_bytes().
// Tile rows are MP_TS uint2 apart (272 bytes): rows 4 apart then start 64 bytes apart modulo 128, so the
// 32 x 16 warp footprint (FOOT 1: 8 column groups x 4 row blocks) is as conflict-free as the 64 x 8 one.
__device__ __forceinline__ uint32_t sq_cap2(uint32_t a, uint32_t b, uint32_t W, uint32_t T)
{   // two x-distances -> packed capped squares
    const uint32_t sa = a > W ? T : a * a, sb = b > W ? T : b * b;
    return sa | (sb << 16);
}

#define LTY_LUT_MAX 8192       // reach LUT in shared memory for T <= this, sqrtf above

// direct: the reach bytes go from registers to global memory (4 bytes per lane and row, 64 contiguous bytes per
// half-warp) instead of through a shared-memory copy of the tile: 16 KB less shared memory, i.e. 4 instead of 3
// resident blocks per SM at W ~ 18
static inline size_t lt_y2_smem_bytes(int Ly, int W, uint32_t T, int direct)
{
    const int rows = ((Ly + 3) & ~3) + 2 * W;
    return (size_t)rows * MP_TS * 8 + 16 + (size_t)(W + 2) * 16 + (direct ? 0 : (size_t)Ly * 128) +
           (T <= LTY_LUT_MAX ? ((T + 16) & ~15u) : 0) + (size_t)((rows + 15) & ~15);
}

template <int FOOT>
__global__ void __launch_bounds__(256)
lt_y2_kernel(const uint8_t *__restrict__ gx, uint8_t *__restrict__ reach, int ny, int nx, uint32_t T,
             int W, int Ly, const int *__restrict__ gate, int direct, const uint8_t *__restrict__ xflag)
{
    if (gate && *gate == 0) return;
    extern __shared__ uint4 lty2_smem[];
    uint2 *tile = reinterpret_cast<uint2 *>(lty2_smem);          // [rows][MP_TS] : 4 x u16 per lane
    const int tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const int rows = ((Ly + 3) & ~3) + 2 * W;                     // Ly rounded up to whole 4-row blocks
    int *range = reinterpret_cast<int *>(tile + (size_t)rows * MP_TS);   // [0] = first useful row, [1] = last
    uint4 *offt = reinterpret_cast<uint4 *>(range + 4);               // [W + 2]: offt[d] = packed capped squares of d .. d+3
    uint32_t *sout = reinterpret_cast<uint32_t *>(offt + (W + 2));    // [Ly][32] reach bytes of the tile (not with `direct`)
    uint8_t *lut = reinterpret_cast<uint8_t *>(sout + (direct ? 0 : (size_t)Ly * 32));   // lut[h] = ceil(sqrt(T - h)), lut[T] = 0
    const bool use_lut = T <= LTY_LUT_MAX;
    uint8_t *rowuse = lut + (use_lut ? ((T + 16) & ~15u) : 0);    // [rows] with xflag: the row holds a value below T
    const int x0 = blockIdx.x * MP_TX, y0 = blockIdx.y * Ly;
    const int64_t zoff = (int64_t)blockIdx.z * ny;
    if (tid == 0) { range[0] = rows; range[1] = -1; }
    if (use_lut)
        for (uint32_t h = tid; h <= T; h += 256) lut[h] = h >= T ? 0 : (uint8_t)ceil_sqrt_small(T - h);
    // offsets are capped at T so that value + offset <= 2T stays inside 16 bits
    for (int d = tid; d < W + 2; d += 256) {
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = min((uint32_t)((d + i) * (d + i)), T) * 0x00010001u;
        offt[d] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // ---- with activity flags from the x pass (one byte per 32-voxel word): which rows of the window can hold a value
    // below T at all; a tile without any writes zeros and leaves, the others skip the loads of their idle rows
    int lo = rows, hi = -1;
    if (xflag) {
        const int nw = nx >> 5, w0 = blockIdx.x * (MP_TX / 32);
        for (int r = tid; r < rows; r += 256) {
            const int y = y0 - W + r;
            uint32_t f = 0;
            if (y >= 0 && y < ny) {
                const uint8_t *fr = xflag + (zoff + y) * nw;
#pragma unroll
                for (int q = 0; q < MP_TX / 32; ++q)
                    if (w0 + q < nw) f |= fr[w0 + q];
            }
            rowuse[r] = f ? 1 : 0;
            if (f) { lo = min(lo, r); hi = max(hi, r); }
        }
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0 && hi >= 0) { atomicMin(&range[0], lo); atomicMax(&range[1], hi); }
    }
    __syncthreads();
    if (xflag && range[1] < 0) {
        for (int i = tid; i < Ly * 8; i += 256) {
            const int r = i >> 3, ch = i & 7;
            const int y = y0 + r, x = x0 + 16 * ch;
            if (y < ny && x < nx) *reinterpret_cast<uint4 *>(reach + (zoff + y) * nx + x) = make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }

    // ---- stage: thread = 16 voxels of one row (8 threads per row, 32 rows per sweep)
    const uint32_t uW = (uint32_t)W;
    lo = rows, hi = -1;
    for (int i0 = 0; i0 < rows * 8; i0 += 4 * 256) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256 + tid, r = i >> 3, ch = i & 7;
            const int y = y0 - W + r, x = x0 + 16 * ch;
            v[u] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (r < rows && y >= 0 && y < ny && x < nx && (!xflag || rowuse[r]))
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
            uint4 *dst = reinterpret_cast<uint4 *>(tile + (size_t)r * MP_TS + 4 * ch);    // row stride 272 = 17 * 16 bytes
            dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
            dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
            if (useful) { lo = min(lo, r); hi = max(hi, r); }
        }
    }
    if (!xflag) {
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0 && hi >= 0) { atomicMin(&range[0], lo); atomicMax(&range[1], hi); }
    }
    __syncthreads();
    const int rlo = range[0], rhi = range[1];

    // ---- scan, register-tiled like edt_minplus_kernel: a lane owns 4 rows x 4 columns; the two
    // rows fetched at step dy serve all 4 outputs.  Warp footprint 64 columns x 8 rows (FOOT 0: 16 column
    // groups x 2 row blocks) or 32 x 16 (FOOT 1: 8 x 4, compacter: the lanes of a warp end their scans closer
    // to each other); a half-warp reads 128 bytes without bank conflicts either way.
    const uint32_t T2 = T * 0x00010001u;
    const int cq = FOOT ? (warp & 3) * 8 + (lane & 7) : (warp & 1) * 16 + (lane & 15);     // uint2 index inside the tile row
    const int ry0 = FOOT ? ((warp >> 2) * 4 + (lane >> 3)) * 4 : ((warp >> 1) * 2 + (lane >> 4)) * 4;
    for (int ry = ry0; ry < Ly; ry += 32) {
        if (y0 + ry >= ny) break;
        const int rr = ry + W;
        uint32_t m[4][4];
        const bool reachable = rhi >= 0 && !(rr + 3 + W < rlo || rr - W > rhi);
        if (!reachable) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) m[i][j] = 0;
        } else {
            uint2 O[4];
            uint32_t B0[4], B1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) O[i] = tile[(rr + i) * MP_TS + cq];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                B0[i] = __vminu2(O[i].x, T2);
                B1[i] = __vminu2(O[i].y, T2);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j != i) {
                        const uint32_t d = (uint32_t)((i - j) * (i - j)) * 0x00010001u;
                        B0[i] = __viaddmin_u16x2(O[j].x, d, B0[i]);
                        B1[i] = __viaddmin_u16x2(O[j].y, d, B1[i]);
                    }
            }
            // rows outside [rlo, rhi] hold nothing below T: clip the scan
            const int dmax = min(W, max(rr + 3 - rlo, rhi - rr));
            // two steps per termination test (a step past the bound only relaxes with valid candidates)
            for (int dy = 1; dy <= dmax; dy += 2) {
                const uint32_t m2 = __vmaxu2(__vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1])),
                                             __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3])));
                const uint32_t bm = max(m2 & 0xFFFFu, m2 >> 16);
                if ((uint32_t)(dy * dy) >= bm) break;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int d = dy + s2;
                    if (d > dmax) break;
                    const uint2 top = tile[(rr - d) * MP_TS + cq];
                    const uint2 bot = tile[(rr + 3 + d) * MP_TS + cq];
                    const uint4 o4 = offt[d];                      // capped (d + i)^2, i = 0..3, both halves
                    const uint32_t of[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        B0[i] = __viaddmin_u16x2(top.x, of[i], B0[i]); B1[i] = __viaddmin_u16x2(top.y, of[i], B1[i]);
                        B0[i] = __viaddmin_u16x2(bot.x, of[3 - i], B0[i]); B1[i] = __viaddmin_u16x2(bot.y, of[3 - i], B1[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t h[4] = {B0[i] & 0xFFFFu, B0[i] >> 16, B1[i] & 0xFFFFu, B1[i] >> 16};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    m[i][j] = use_lut ? (uint32_t)lut[min(h[j], T)] : (h[j] >= T ? 0u : ceil_sqrt_small(T - h[j]));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (ry + i < Ly) {
                const uint32_t pk = pack4(m[i][0], m[i][1], m[i][2], m[i][3]);
                if (!direct) sout[(ry + i) * 32 + cq] = pk;
                else if (y0 + ry + i < ny && x0 + 4 * cq < nx)
                    *reinterpret_cast<uint32_t *>(reach + (zoff + y0 + ry + i) * nx + x0 + 4 * cq) = pk;
            }
    }
    if (direct) return;
    __syncthreads();
    // ---- coalesced write-out of the reach tile (16 bytes per thread)
    for (int i = tid; i < Ly * 8; i += 256) {
        const int r = i >> 3, ch = i & 7;
        const int y = y0 + r, x = x0 + 16 * ch;
        if (y < ny && x < nx)
            *reinterpret_cast<uint4 *>(reach + (zoff + y) * nx + x) = reinterpret_cast<const uint4 *>(sout)[i];
    }
}

// ------------------------------------------- per-radius y pass, hierarchical scan (the default)
// Same tile, same result as lt_y2_kernel.  What bounds lt_y2 is not the scan of the voxels that find a seed
// (it ends after about sqrt(h) steps) but the voxels that find none: they walk all W rows, nearly every
// 64 x 8 warp patch holds one, and a warp is as slow as its slowest lane.  Here the tile also carries, for
// every aligned group of 4 rows and every 4-column group, the minimum of its 16 values (`cm`); the steps
// dy = 4g-3 .. 4g of a block read exactly the row groups g above and below it, so
//       min(cm_above, cm_below) + (4g - 3)^2  >=  max(best of the block)
// proves that those four steps cannot change anything and they are skipped (two 2-byte loads and a compare
// instead of 8 row loads and 64 VIADDMNMX).  The decision is taken per warp (`__any_sync`), so there is no
// divergence: a group is scanned by all lanes as soon as one lane needs it, which is harmless (a relaxation
// with a valid candidate never hurts).  Halo = W rounded up to a multiple of 4 rows so that groups align.
static inline size_t lt_y3_smem_bytes(int Ly, int W, uint32_t T, int direct)
{
    const int Hh = (W + 3) & ~3, rows = ((Ly + 3) & ~3) + 2 * Hh;
    return (size_t)rows * MP_TS * 8 + (size_t)rows * 16 + 16 + (size_t)(Hh + 6) * 16 + (direct ? 0 : (size_t)Ly * 128) +
           (T <= LTY_LUT_MAX ? ((T + 16) & ~15u) : 0);
}

__global__ void __launch_bounds__(256)
lt_y3_kernel(const uint8_t *__restrict__ gx, uint8_t *__restrict__ reach, int ny, int nx, uint32_t T,
             int W, int Ly, const int *__restrict__ gate, int direct)
{
    if (gate && *gate == 0) return;
    extern __shared__ uint4 lty3_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const int Hh = (W + 3) & ~3;                                  // halo rows: whole groups of 4
    const int rows = ((Ly + 3) & ~3) + 2 * Hh, ngroups = rows >> 2;
    uint2 *tile = reinterpret_cast<uint2 *>(lty3_smem);          // [rows][MP_TS] : 4 x u16 per entry
    uint16_t *cm = reinterpret_cast<uint16_t *>(tile + (size_t)rows * MP_TS);   // [rows / 4][32] group minima
    int *range = reinterpret_cast<int *>(cm + (size_t)ngroups * 32);            // [0] first useful row, [1] last
    uint4 *offt = reinterpret_cast<uint4 *>(range + 4);          // [Hh + 6]: packed capped squares of d .. d+3
    uint32_t *sout = reinterpret_cast<uint32_t *>(offt + (Hh + 6));             // [Ly][32] reach bytes of the tile (not with `direct`)
    uint8_t *lut = reinterpret_cast<uint8_t *>(sout + (direct ? 0 : (size_t)Ly * 32));
    const bool use_lut = T <= LTY_LUT_MAX;
    const int x0 = blockIdx.x * MP_TX, y0 = blockIdx.y * Ly;
    const int64_t zoff = (int64_t)blockIdx.z * ny;
    if (tid == 0) { range[0] = rows; range[1] = -1; }
    if (use_lut)
        for (uint32_t h = tid; h <= T; h += 256) lut[h] = h >= T ? 0 : (uint8_t)ceil_sqrt_small(T - h);
    for (int d = tid; d < Hh + 6; d += 256) {
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = min((uint32_t)((d + i) * (d + i)), T) * 0x00010001u;
        offt[d] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();

    // ---- stage: thread = 16 voxels of one row (8 threads per row, 32 rows per sweep)
    const uint32_t uW = (uint32_t)W;
    int lo = rows, hi = -1;
    for (int i0 = 0; i0 < rows * 8; i0 += 4 * 256) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256 + tid, r = i >> 3, ch = i & 7;
            const int y = y0 - Hh + r, x = x0 + 16 * ch;
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
            uint4 *dst = reinterpret_cast<uint4 *>(tile + (size_t)r * MP_TS + 4 * ch);
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
    // ---- group minima (skipped when the tile holds no seed at all: every output is 0 then)
    if (rhi >= 0)
        for (int i = tid; i < ngroups * 32; i += 256) {
            const int q = i >> 5, c = i & 31;
            const uint2 a = tile[(4 * q) * MP_TS + c], b = tile[(4 * q + 1) * MP_TS + c];
            const uint2 e = tile[(4 * q + 2) * MP_TS + c], f = tile[(4 * q + 3) * MP_TS + c];
            const uint32_t m2 = __vminu2(__vminu2(__vminu2(a.x, a.y), __vminu2(b.x, b.y)),
                                         __vminu2(__vminu2(e.x, e.y), __vminu2(f.x, f.y)));
            cm[i] = (uint16_t)min(m2 & 0xFFFFu, m2 >> 16);
        }
    __syncthreads();

    // ---- scan: a lane owns 4 rows x 4 columns, warp footprint 64 columns x 8 rows (as lt_y2_kernel)
    const uint32_t T2 = T * 0x00010001u;
    const int cq = (warp & 1) * 16 + (lane & 15);
    const int ry0 = ((warp >> 1) * 2 + (lane >> 4)) * 4;
    const int iters = (Ly + 31) >> 5;                            // the same for every lane: the loop holds warp votes
    for (int t = 0; t < iters; ++t) {
        const int ry = ry0 + 32 * t;
        const bool valid = ry < Ly && y0 + ry < ny;
        const int rr = min(ry, ((Ly + 3) & ~3) - 4) + Hh;        // (lanes past the end scan a valid block, unused)
        const int qb = rr >> 2;
        const bool reachable = valid && rhi >= 0 && !(rr + 3 + W < rlo || rr - W > rhi);
        uint2 O[4];
        uint32_t B0[4], B1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) O[i] = tile[(rr + i) * MP_TS + cq];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            B0[i] = __vminu2(O[i].x, T2);
            B1[i] = __vminu2(O[i].y, T2);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j != i) {
                    const uint32_t d = (uint32_t)((i - j) * (i - j)) * 0x00010001u;
                    B0[i] = __viaddmin_u16x2(O[j].x, d, B0[i]);
                    B1[i] = __viaddmin_u16x2(O[j].y, d, B1[i]);
                }
        }
        // rows outside [rlo, rhi] hold nothing below T: clip the scan
        const int dmax = reachable ? min(W, max(rr + 3 - rlo, rhi - rr)) : 0;
        uint32_t m2 = __vmaxu2(__vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1])),
                               __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3])));
        uint32_t bm = max(m2 & 0xFFFFu, m2 >> 16);
        for (int g = 1; 4 * g <= Hh; ++g) {
            const int dlo = 4 * g - 3;
            const bool active = dlo <= dmax && (uint32_t)(dlo * dlo) < bm;
            if (!__any_sync(0xFFFFFFFFu, active)) break;
            const uint32_t cu = cm[(qb - g) * 32 + cq], cd = cm[(qb + g) * 32 + cq];
            const bool need = active && min(cu, cd) + (uint32_t)(dlo * dlo) < bm;
            if (!__any_sync(0xFFFFFFFFu, need)) continue;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int d = dlo + s4;
                const uint2 top = tile[(rr - d) * MP_TS + cq];
                const uint2 bot = tile[(rr + 3 + d) * MP_TS + cq];
                const uint4 o4 = offt[d];                          // capped (d + i)^2, i = 0..3, both halves
                const uint32_t of[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    B0[i] = __viaddmin_u16x2(top.x, of[i], B0[i]); B1[i] = __viaddmin_u16x2(top.y, of[i], B1[i]);
                    B0[i] = __viaddmin_u16x2(bot.x, of[3 - i], B0[i]); B1[i] = __viaddmin_u16x2(bot.y, of[3 - i], B1[i]);
                }
            }
            m2 = __vmaxu2(__vmaxu2(__vmaxu2(B0[0], B1[0]), __vmaxu2(B0[1], B1[1])),
                          __vmaxu2(__vmaxu2(B0[2], B1[2]), __vmaxu2(B0[3], B1[3])));
            bm = max(m2 & 0xFFFFu, m2 >> 16);
        }
        if (!valid) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t m[4];
            const uint32_t h[4] = {B0[i] & 0xFFFFu, B0[i] >> 16, B1[i] & 0xFFFFu, B1[i] >> 16};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                m[j] = !reachable ? 0u : (use_lut ? (uint32_t)lut[min(h[j], T)] : (h[j] >= T ? 0u : ceil_sqrt_small(T - h[j])));
            if (ry + i < Ly) {
                const uint32_t pk = pack4(m[0], m[1], m[2], m[3]);
                if (!direct) sout[(ry + i) * 32 + cq] = pk;
                else if (y0 + ry + i < ny && x0 + 4 * cq < nx)
                    *reinterpret_cast<uint32_t *>(reach + (zoff + y0 + ry + i) * nx + x0 + 4 * cq) = pk;
            }
        }
    }
    if (direct) return;
    __syncthreads();
    // ---- coalesced write-out of the reach tile (16 bytes per thread)
    for (int i = tid; i < Ly * 8; i += 256) {
        const int r = i >> 3, ch = i & 7;
        const int y = y0 + r, x = x0 + 16 * ch;
        if (y < ny && x < nx)
            *reinterpret_cast<uint4 *>(reach + (zoff + y) * nx + x) = reinterpret_cast<const uint4 *>(sout)[i];
    }
}
