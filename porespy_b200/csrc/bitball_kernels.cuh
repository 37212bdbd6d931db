// bitball_kernels.cuh -- bit-parallel sphere insertion for small radii.
//
// For a threshold T the per-radius step of porosimetry / local_thickness
// (/root/reference/src/porespy/filters/_funcs.py:1180-1192 and :1196-1209) is
//     fill = dilation of the seed set by the digital ball  B_T = {o : |o|^2 < T}
// (edt(~seeds) < r  ==  fftconvolve(seeds, ps_ball(r)) > 0.1).  With one
// bit per voxel (32 voxels per word along x) that dilation is
//     fill(y,z) = OR_{a = 0..W}  dil_x^a ( OR_{(dy,dz) in ring_a} seeds(y+dy, z+dz) )
// where ring_a holds the (dy,dz) whose x-allowance max{|dx| : dx^2+dy^2+dz^2 < T} is exactly a
// and dil_x^a is a 1-voxel dilation along x applied a times (Horner scheme from a = W down
// to 0).  Cost: about pi*T word-ORs plus 6*W shuffle/shift ops per 32 voxels, i.e. ~0.1
// instructions per voxel for T = 4 and ~3 for T = 25, against ~100 per voxel for the
// byte pipeline -- the per-voxel work of the byte kernels, not HBM, is what bounds them.
//
// A warp owns one output row segment of 32 words (1024 voxels), a lane one word; source rows
// are read straight from global memory (a 128-byte coalesced line per (dy,dz), L1-resident
// across the 8 rows of a block).  Rows longer than 1024 voxels use 30-word segments with one
// halo word on each side.
#pragma once
#include "common.cuh"
#include "xdist_kernels.cuh"

#define BB_MAX_PAIRS 1408         // pi * 400 pairs + every ring padded to a multiple of 4
#define BB_TY 8                    // output rows of a block: 8 (y) x 4 (z) -- one warp each
#define BB_TZ 4
// Every ring is padded to a multiple of 4 pairs with copies of its last pair (OR is idempotent), so
// the kernels fetch the word offsets of 4 pairs with one 16-byte shared-memory load: the offsets
// live in the kernel parameter (constant bank), and an indexed LDC per pair was 40 % of the
// stall samples of the dilation kernel.
struct BallPairs {                 // (dy,dz) offsets sorted by x-allowance a, descending
    int W;                         // largest allowance = ceil(sqrt(T)) - 1
    unsigned short ring_end[34];   // pairs of ring a are [ring_end[a + 1], ring_end[a]) for a = W..0
    int2 e[BB_MAX_PAIRS];          // .x = (dz * ny + dy) * nw (word offset), .y = (dy & 0xFFFF) | (dz << 16)
};

// ------------------------------------------------------------------ class map -> seed bits
// bits[row][w] bit b = (cls[row][32 w + b] <= k).  nx % 32 == 0.  One thread per word.
__global__ void __launch_bounds__(256)
lt_pack_kernel(const uint8_t *__restrict__ cls, uint32_t *__restrict__ bits, int64_t nwords, int k,
               const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const uint32_t n = (uint32_t)(k + 1);                 // non-seed <=> byte >= n
    const uint32_t nl4 = (n & 0x7Fu) * 0x01010101u;
    const uint32_t sel = n < 128u ? 0xFFFFFFFFu : 0u;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += step) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(cls) + 2 * w);
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(cls) + 2 * w + 1);
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t m = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) m |= gather_bit7(swar_ge(v[q], nl4, sel) ^ 0x80808080u) << (4 * q);
        bits[w] = m;
    }
}

// Up to PACKN_MAX consecutive radii from ONE read of the class map, bit-sliced: the class bytes of a word's 32
// voxels are transposed into bit planes (nb low planes + one "any higher bit" mask), and  cls <= k  is then a
// handful of bitwise operations per radius on whole words (a most-significant-bit-first comparator) instead
// of a byte-SWAR compare per radius and per 4 voxels.  bits of radius k0 + i go to bits + i * vol_words.
#define PACKN_MAX 16
// NB = number of low bit planes the comparator looks at (the largest radius index of the call is < 2^NB);
// a compile-time constant so that the planes stay in registers and the loops unroll without predicates.
template <int NB>
__global__ void __launch_bounds__(256)
lt_packn_kernel(const uint8_t *__restrict__ cls, uint32_t *__restrict__ bits, int64_t nwords, int64_t vol_words,
                int k0, int nk)
{
    constexpr uint32_t hn = NB < 8 ? (1u << NB) : 0u;                   // "high" bytes are >= 2^NB
    constexpr uint32_t hl4 = (hn & 0x7Fu) * 0x01010101u, hsel = hn < 128u ? 0xFFFFFFFFu : 0u;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += step) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(cls) + 2 * w);
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(cls) + 2 * w + 1);
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t p[NB], hi = 0u;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            uint32_t pj = 0u;
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                // bit j of the 8 bytes of (v[q], v[q+1]) -> one byte: the multiply moves the bits at positions
                // 0,8,16,24 (v[q]) and 4,12,20,28 (v[q+1], pre-shifted by 4) to 24..31 without carries
                const uint32_t t = ((v[q] >> j) & 0x01010101u) | (((v[q + 1] >> j) & 0x01010101u) << 4);
                pj |= ((t * 0x01020408u) >> 24) << (4 * q);
            }
            p[j] = pj;
        }
        if (NB < 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) hi |= gather_bit7(swar_ge(v[q], hl4, hsel)) << (4 * q);
        }
        for (int i = 0; i < nk; ++i) {
            const uint32_t k = (uint32_t)(k0 + i);
            uint32_t lt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
            for (int j = NB - 1; j >= 0; --j) {
                const uint32_t m = 0u - ((k >> j) & 1u);                 // all ones where bit j of k is set (uniform)
                lt |= eq & ~p[j] & m;
                eq &= ~(p[j] ^ m);
            }
            bits[(int64_t)i * vol_words + w] = (lt | eq) & ~hi;
        }
    }
}

// written[row][w] bit b = (idx[row][32 w + b] != 0)
__global__ void __launch_bounds__(256)
lt_wmask_kernel(const uint8_t *__restrict__ idx, uint32_t *__restrict__ written, int64_t nwords)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += step) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(idx) + 2 * w);
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(idx) + 2 * w + 1);
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t m = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) m |= gather_bit7(swar_ge(v[q], 0x01010101u, 0xFFFFFFFFu)) << (4 * q);
        written[w] = m;
    }
}

// New voxels of one word: set their written bits and their radius-index bytes.
__device__ __forceinline__ void bb_commit(uint32_t *__restrict__ written, uint8_t *__restrict__ idx, int64_t wi,
                                          uint32_t A, uint32_t val)
{
    const uint32_t wm = written[wi];
    const uint32_t N = A & ~wm;
    if (N == 0u) return;
    written[wi] = wm | N;
    uint4 *ip = reinterpret_cast<uint4 *>(idx + 32 * wi);
    const uint32_t val4 = val * 0x01010101u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t nh = (N >> (16 * h)) & 0xFFFFu;
        if (nh == 0u) continue;
        uint4 o = ip[h];
        uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t nib = (nh >> (4 * q)) & 0xFu;
            const uint32_t bm = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;   // nibble -> byte mask
            ow[q] |= val4 & bm;        // the bytes under bm are still 0 (written bit was clear)
        }
        ip[h] = o;
    }
}

// Four consecutive words (wi a multiple of 4: 16-byte aligned in `written`, 128-byte aligned in `idx`): the written
// words arrive as one load, and the index halves that receive new voxels are fetched two words (up to four 16-byte
// loads) at a time before any of them is stored -- word by word every load waits behind the previous word's stores
// (same arrays: the compiler must keep the order), a dozen serial memory round trips at the end of every thread.
__device__ __forceinline__ void bb_commit4(uint32_t *__restrict__ written, uint8_t *__restrict__ idx, int64_t wi,
                                           uint32_t A0, uint32_t A1, uint32_t A2, uint32_t A3, uint32_t val)
{
    const uint4 wm = *reinterpret_cast<const uint4 *>(written + wi);
    const uint32_t N[4] = {A0 & ~wm.x, A1 & ~wm.y, A2 & ~wm.z, A3 & ~wm.w};
    if ((N[0] | N[1] | N[2] | N[3]) == 0u) return;
    *reinterpret_cast<uint4 *>(written + wi) = make_uint4(wm.x | N[0], wm.y | N[1], wm.z | N[2], wm.w | N[3]);
    uint4 *ip = reinterpret_cast<uint4 *>(idx + 32 * wi);              // two 16-byte halves per word
    const uint32_t val4 = val * 0x01010101u;
#pragma unroll
    for (int g = 0; g < 2; ++g) {                                       // words 2g, 2g + 1
        if ((N[2 * g] | N[2 * g + 1]) == 0u) continue;
        uint4 o[4];
        uint32_t nh[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            nh[t] = (N[2 * g + (t >> 1)] >> (16 * (t & 1))) & 0xFFFFu;
            if (nh[t]) o[t] = ip[4 * g + t];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (nh[t] == 0u) continue;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&o[t]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t nib = (nh[t] >> (4 * q)) & 0xFu;
                const uint32_t bm = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;   // nibble -> byte mask
                ow[q] |= val4 & bm;        // the bytes under bm are still 0 (written bit was clear)
            }
            ip[4 * g + t] = o[t];
        }
    }
}

// ------------------------------------------------------------------------------ dilation
// seeds: [nz_src][ny][nw] words (nw = nx / 32); output plane z reads seed plane z + z_off, so a
// z-slab shard passes its slab with the neighbours' halo planes in front / behind (single GPU:
// nz_src = nz, z_off = 0).  written: [nz][ny][nw].  idx: [nz][ny][nx] bytes.
// grid = (nseg, ceil(ny / 8), ceil(nz / 4)), block = 1024 (8 x 4 rows: the source rows of a
// block, (8 + 2W) x (4 + 2W) x 128 bytes, stay in L1).  seg_words = 32 (nw <= 32) or 30.
__global__ void __launch_bounds__(1024)
lt_bitball_kernel(const uint32_t *__restrict__ seeds, uint32_t *__restrict__ written,
                  uint8_t *__restrict__ idx, int nz, int ny, int nw, int seg_words,
                  const __grid_constant__ BallPairs bp, uint32_t val, const int *__restrict__ gate,
                  int nz_src, int z_off)
{
    if (gate && *gate == 0) return;
    __shared__ __align__(16) int s_off[BB_MAX_PAIRS];
    for (int i = threadIdx.x; i < (int)bp.ring_end[0]; i += blockDim.x) s_off[i] = bp.e[i].x;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int y = blockIdx.y * BB_TY + (warp & (BB_TY - 1)), z = blockIdx.z * BB_TZ + (warp >> 3);
    if (y >= ny || z >= nz) return;
    // word owned by this lane; with 30-word segments lanes 0 and 31 are halo words
    const int halo = seg_words == 32 ? 0 : 1;
    const int w = blockIdx.x * seg_words + lane - halo;
    const bool inrow = w >= 0 && w < nw;
    const uint32_t inmask = inrow ? 0xFFFFFFFFu : 0u;
    const int zs = z + z_off;             // plane of this row inside the seed buffer (z halo planes first)
    const uint32_t *base = seeds + ((int64_t)zs * ny + y) * nw + (inrow ? w : 0);
    asm volatile("" : "+l"(base));        // keep it one pointer: base + offset is then a single IMAD.WIDE
    const int W = bp.W;
    const bool interior = y - W >= 0 && y + W < ny && zs - W >= 0 && zs + W < nz_src;

    uint32_t A = 0;
    int p = 0;
    for (int a = W; a >= 0; --a) {
        if (a < W) {
            uint32_t l = __shfl_up_sync(0xFFFFFFFFu, A, 1), r = __shfl_down_sync(0xFFFFFFFFu, A, 1);
            if (lane == 0) l = 0;
            if (lane == 31) r = 0;
            A |= (A << 1) | (l >> 31) | (A >> 1) | (r << 31);
        }
        const int pend = bp.ring_end[a];
        if (interior) {
            for (; p < pend; p += 4) {                 // rings are padded to whole groups of 4
                const int4 o = *reinterpret_cast<const int4 *>(s_off + p);
                const uint32_t s0 = __ldg(base + o.x), s1 = __ldg(base + o.y);
                const uint32_t s2 = __ldg(base + o.z), s3 = __ldg(base + o.w);
                A |= ((s0 | s1) | (s2 | s3)) & inmask;
            }
        } else {
            for (; p < pend; ++p) {
                const int2 e = bp.e[p];
                const int yy = y + (int)(short)(e.y & 0xFFFF), zz = zs + (e.y >> 16);
                if ((unsigned)yy < (unsigned)ny && (unsigned)zz < (unsigned)nz_src) A |= __ldg(base + e.x) & inmask;
            }
        }
    }
    const bool center = inrow && (halo == 0 || (lane >= 1 && lane <= 30));
    if (!center) return;
    bb_commit(written, idx, ((int64_t)z * ny + y) * nw + w, A, val);
}

// Two words per lane: a warp owns 64 words (2048 voxels) of a row, so rows of 33..64 words need no
// halo lanes (with 32-word warps a 2048-voxel row costs three 30-word segments, a third of the
// lanes idle); longer rows use 60-word segments with one halo lane (two words) on each side.
// nw must be even (8-byte loads).  grid.x = segments, otherwise as lt_bitball_kernel.
__global__ void __launch_bounds__(1024)
lt_bitball2_kernel(const uint32_t *__restrict__ seeds, uint32_t *__restrict__ written,
                   uint8_t *__restrict__ idx, int nz, int ny, int nw, int seg_words,
                   const __grid_constant__ BallPairs bp, uint32_t val, const int *__restrict__ gate,
                   int nz_src, int z_off)
{
    if (gate && *gate == 0) return;
    __shared__ __align__(16) int s_off[BB_MAX_PAIRS];
    for (int i = threadIdx.x; i < (int)bp.ring_end[0]; i += blockDim.x) s_off[i] = bp.e[i].x;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int y = blockIdx.y * BB_TY + (warp & (BB_TY - 1)), z = blockIdx.z * BB_TZ + (warp >> 3);
    if (y >= ny || z >= nz) return;
    const int halo = seg_words == 64 ? 0 : 2;
    const int w = blockIdx.x * seg_words + 2 * lane - halo;      // first of this lane's two words (even)
    const bool inrow = w >= 0 && w < nw;
    const uint32_t inmask = inrow ? 0xFFFFFFFFu : 0u;
    const int zs = z + z_off;
    const uint32_t *base = seeds + ((int64_t)zs * ny + y) * nw + (inrow ? w : 0);
    asm volatile("" : "+l"(base));
    const int W = bp.W;
    const bool interior = y - W >= 0 && y + W < ny && zs - W >= 0 && zs + W < nz_src;

    uint32_t A0 = 0, A1 = 0;
    int p = 0;
    for (int a = W; a >= 0; --a) {
        if (a < W) {
            uint32_t l = __shfl_up_sync(0xFFFFFFFFu, A1, 1), r = __shfl_down_sync(0xFFFFFFFFu, A0, 1);
            if (lane == 0) l = 0;
            if (lane == 31) r = 0;
            const uint32_t n0 = A0 | (A0 << 1) | (l >> 31) | (A0 >> 1) | (A1 << 31);
            const uint32_t n1 = A1 | (A1 << 1) | (A0 >> 31) | (A1 >> 1) | (r << 31);
            A0 = n0;
            A1 = n1;
        }
        const int pend = bp.ring_end[a];
        if (interior) {
            for (; p < pend; p += 4) {                 // rings are padded to whole groups of 4
                const int4 o = *reinterpret_cast<const int4 *>(s_off + p);
                const uint2 s0 = __ldg(reinterpret_cast<const uint2 *>(base + o.x));
                const uint2 s1 = __ldg(reinterpret_cast<const uint2 *>(base + o.y));
                const uint2 s2 = __ldg(reinterpret_cast<const uint2 *>(base + o.z));
                const uint2 s3 = __ldg(reinterpret_cast<const uint2 *>(base + o.w));
                A0 |= ((s0.x | s1.x) | (s2.x | s3.x)) & inmask;
                A1 |= ((s0.y | s1.y) | (s2.y | s3.y)) & inmask;
            }
        } else {
            for (; p < pend; ++p) {
                const int2 e = bp.e[p];
                const int yy = y + (int)(short)(e.y & 0xFFFF), zz = zs + (e.y >> 16);
                if ((unsigned)yy < (unsigned)ny && (unsigned)zz < (unsigned)nz_src) {
                    const uint2 s0 = __ldg(reinterpret_cast<const uint2 *>(base + e.x));
                    A0 |= s0.x & inmask;
                    A1 |= s0.y & inmask;
                }
            }
        }
    }
    const bool center = inrow && (halo == 0 || (lane >= 1 && lane <= 30));
    if (!center) return;
    const int64_t wi = ((int64_t)z * ny + y) * nw + w;
    bb_commit(written, idx, wi, A0, val);
    bb_commit(written, idx, wi + 1, A1, val);
}

// Four words per lane (one 16-byte load per (dy,dz) pair and lane): a row of nw = 4 * LPR words is
// owned by LPR lanes, a warp owns 32 / LPR rows.  Per word this issues half the instructions of
// the one-word kernel (the OR work is the same, the address arithmetic, the load and the loop
// overhead are shared by four words) for the same L1 wavefronts per byte.
//   LPR = 8 (nx = 1024): warp = 4 rows, block (256 threads) = 8 (y) x 4 (z) rows
//   LPR = 16 (nx = 2048): warp = 2 rows, block = 8 x 2;   LPR = 32 (nx = 4096): warp = 1 row, block = 8 x 1
// grid = (1, ceil(ny / 8), ceil(nz / TZ)), TZ = 32 / LPR.  seeds / written 16-byte aligned.
template <int LPR>
__global__ void __launch_bounds__(256, 6)
lt_bitball4_kernel(const uint32_t *__restrict__ seeds, uint32_t *__restrict__ written,
                   uint8_t *__restrict__ idx, int nz, int ny,
                   const __grid_constant__ BallPairs bp, uint32_t val, const int *__restrict__ gate,
                   int nz_src, int z_off)
{
    if (gate && *gate == 0) return;
    constexpr int RPW = 32 / LPR;          // rows per warp
    constexpr int NW = 4 * LPR;            // words per row
    __shared__ __align__(16) int s_off[BB_MAX_PAIRS];
    __shared__ int s_dyz[BB_MAX_PAIRS];
    for (int i = threadIdx.x; i < (int)bp.ring_end[0]; i += blockDim.x) {
        s_off[i] = bp.e[i].x;
        s_dyz[i] = bp.e[i].y;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int r = warp * RPW + lane / LPR;                 // row of the block: 8 (y) x RPW (z)
    const int lr = lane % LPR;                             // lane inside the row
    const int y = blockIdx.y * 8 + (r & 7), z = blockIdx.z * RPW + (r >> 3);
    const bool rowok = y < ny && z < nz;
    const uint32_t inmask = rowok ? 0xFFFFFFFFu : 0u;
    const int yc = rowok ? y : 0, zc = rowok ? z : 0;      // rows past the end read row (0, 0) and are masked
    const int zs = zc + z_off;
    const uint32_t *base = seeds + ((int64_t)zs * ny + yc) * NW + 4 * lr;
    asm volatile("" : "+l"(base));
    const int W = bp.W;
    const bool interior = yc - W >= 0 && yc + W < ny && zs - W >= 0 && zs + W < nz_src;

    uint32_t A0 = 0, A1 = 0, A2 = 0, A3 = 0;
    int p = 0;
    for (int a = W; a >= 0; --a) {
        if (a < W) {
            uint32_t l = __shfl_up_sync(0xFFFFFFFFu, A3, 1), rr = __shfl_down_sync(0xFFFFFFFFu, A0, 1);
            if (lr == 0) l = 0;                            // row boundary (also a warp-internal one)
            if (lr == LPR - 1) rr = 0;
            const uint32_t n0 = A0 | (A0 << 1) | (l >> 31) | (A0 >> 1) | (A1 << 31);
            const uint32_t n1 = A1 | (A1 << 1) | (A0 >> 31) | (A1 >> 1) | (A2 << 31);
            const uint32_t n2 = A2 | (A2 << 1) | (A1 >> 31) | (A2 >> 1) | (A3 << 31);
            const uint32_t n3 = A3 | (A3 << 1) | (A2 >> 31) | (A3 >> 1) | (rr << 31);
            A0 = n0; A1 = n1; A2 = n2; A3 = n3;
        }
        const int pend = bp.ring_end[a];
        if (interior) {
            for (; p < pend; p += 4) {                     // rings are padded to whole groups of 4
                const int4 o = *reinterpret_cast<const int4 *>(s_off + p);
                const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + o.x));
                const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(base + o.y));
                const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(base + o.z));
                const uint4 s3 = __ldg(reinterpret_cast<const uint4 *>(base + o.w));
                A0 |= ((s0.x | s1.x) | (s2.x | s3.x)) & inmask;
                A1 |= ((s0.y | s1.y) | (s2.y | s3.y)) & inmask;
                A2 |= ((s0.z | s1.z) | (s2.z | s3.z)) & inmask;
                A3 |= ((s0.w | s1.w) | (s2.w | s3.w)) & inmask;
            }
        } else {
            for (; p < pend; ++p) {
                const int e = s_dyz[p];
                const int yy = yc + (int)(short)(e & 0xFFFF), zz = zs + (e >> 16);
                if ((unsigned)yy < (unsigned)ny && (unsigned)zz < (unsigned)nz_src) {
                    const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + s_off[p]));
                    A0 |= s0.x & inmask; A1 |= s0.y & inmask; A2 |= s0.z & inmask; A3 |= s0.w & inmask;
                }
            }
        }
    }
    if (!rowok) return;
    const int64_t wi = ((int64_t)z * ny + y) * NW + 4 * lr;
    bb_commit4(written, idx, wi, A0, A1, A2, A3, val);
}

// ------------------------------------------------- four words x two output rows per lane
// The dilation is bound by L1 wavefronts: one 16-byte load per (dy, dz) pair and output row.  A source row
// (y + sy, z + sz) serves the outputs (y, z) and (y + 1, z) with the allowances allow(sy, sz) and allow(sy - 1, sz);
// where the two coincide the row is loaded ONCE in that Horner stage and ORed into both accumulators.  For the
// balls of this path that needs 79 % of the loads per output (CPU count; T = 180: 884 loads for 2 x 561 pairs).
// Every stage has three lists -- rows for both outputs, for the upper one only, for the lower one only -- each
// padded to a multiple of 4 so that the loops keep the 4-loads-per-iteration form of lt_bitball4_kernel.
// Measured and NOT adopted (r2k): 2 x 2 outputs with a 4-bit output mask per load (58 % of the loads): the masked
// ORs double the instruction count and the kernel turns issue-bound (3.26 -> 3.78 ms at T = 180).
// Row groups whose two rows touch the volume border take the one-output loop with bounds checks.
//   LPR lanes own one row group (4 * LPR words per row); a block of 256 threads owns 256 / LPR row groups:
//   8 (y pairs) x (32 / LPR) (z).  grid = (1, ceil(ny / 16), ceil(nz / (32 / LPR))).
#define BB3_MAX_ENTRIES 2432
struct BallDuo {                   // word offsets sorted by stage (x-allowance a, descending) and list
    int W;
    unsigned short end[34][3];     // stage a: both = [prev, end[a][0]), upper only = [end[a][0], end[a][1]), lower only = [.., end[a][2])
    int off[BB3_MAX_ENTRIES];      // (sz * ny + sy) * nw, relative to the upper output row
};

template <int LPR>
__global__ void __launch_bounds__(256, 4)
lt_bitball4d_kernel(const uint32_t *__restrict__ seeds, uint32_t *__restrict__ written,
                    uint8_t *__restrict__ idx, int nz, int ny,
                    const __grid_constant__ BallDuo bd, const __grid_constant__ BallPairs bp, uint32_t val,
                    const int *__restrict__ gate, int nz_src, int z_off)
{
    if (gate && *gate == 0) return;
    constexpr int GPB = 256 / LPR;         // row groups per block
    constexpr int GZ = GPB / 8;            // ... of which along z (8 along y)
    constexpr int NW = 4 * LPR;            // words per row
    __shared__ __align__(16) int s_d[BB3_MAX_ENTRIES];
    __shared__ int s_off[BB_MAX_PAIRS];
    __shared__ int s_dyz[BB_MAX_PAIRS];
    for (int i = threadIdx.x; i < (int)bd.end[0][2]; i += blockDim.x) s_d[i] = bd.off[i];
    for (int i = threadIdx.x; i < (int)bp.ring_end[0]; i += blockDim.x) { s_off[i] = bp.e[i].x; s_dyz[i] = bp.e[i].y; }
    __syncthreads();
    const int grp = threadIdx.x / LPR, lr = threadIdx.x % LPR;
    const int y0 = (blockIdx.y * 8 + (grp & 7)) * 2, z = blockIdx.z * GZ + (grp >> 3);
    if (y0 >= ny || z >= nz) return;
    const int W = bd.W;
    const int zs = z + z_off;
    const bool interior = y0 - W >= 0 && y0 + 1 + W < ny && zs - W >= 0 && zs + W < nz_src;
    // the lanes of one row group take the same branches; a warp may hold several groups, so every shuffle names
    // only the lanes of its own group
    const unsigned gmask = LPR == 32 ? 0xFFFFFFFFu : (((1u << LPR) - 1u) << ((lane_id() / LPR) * LPR));
#define BB_DILATE4(P0, P1, P2, P3)                                                                        \
    {                                                                                                     \
        uint32_t l = __shfl_up_sync(gmask, P3, 1), rr = __shfl_down_sync(gmask, P0, 1);                   \
        if (lr == 0) l = 0;                                                                               \
        if (lr == LPR - 1) rr = 0;                                                                        \
        const uint32_t n0 = P0 | (P0 << 1) | (l >> 31) | (P0 >> 1) | (P1 << 31);                          \
        const uint32_t n1 = P1 | (P1 << 1) | (P0 >> 31) | (P1 >> 1) | (P2 << 31);                         \
        const uint32_t n2 = P2 | (P2 << 1) | (P1 >> 31) | (P2 >> 1) | (P3 << 31);                         \
        const uint32_t n3 = P3 | (P3 << 1) | (P2 >> 31) | (P3 >> 1) | (rr << 31);                         \
        P0 = n0; P1 = n1; P2 = n2; P3 = n3;                                                               \
    }
    if (interior) {
        const uint32_t *base = seeds + ((int64_t)zs * ny + y0) * NW + 4 * lr;
        asm volatile("" : "+l"(base));
        uint32_t A0 = 0, A1 = 0, A2 = 0, A3 = 0, B0 = 0, B1 = 0, B2 = 0, B3 = 0;
        int p = 0;
        for (int a = W; a >= 0; --a) {
            if (a < W) {
                BB_DILATE4(A0, A1, A2, A3)
                BB_DILATE4(B0, B1, B2, B3)
            }
            const int e0 = bd.end[a][0], e1 = bd.end[a][1], e2 = bd.end[a][2];
            for (; p < e0; p += 4) {                        // rows that serve both outputs
                const int4 o = *reinterpret_cast<const int4 *>(s_d + p);
                const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + o.x));
                const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(base + o.y));
                const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(base + o.z));
                const uint4 s3 = __ldg(reinterpret_cast<const uint4 *>(base + o.w));
                const uint32_t c0 = (s0.x | s1.x) | (s2.x | s3.x), c1 = (s0.y | s1.y) | (s2.y | s3.y);
                const uint32_t c2 = (s0.z | s1.z) | (s2.z | s3.z), c3 = (s0.w | s1.w) | (s2.w | s3.w);
                A0 |= c0; A1 |= c1; A2 |= c2; A3 |= c3;
                B0 |= c0; B1 |= c1; B2 |= c2; B3 |= c3;
            }
            for (; p < e1; p += 4) {                        // upper output only
                const int4 o = *reinterpret_cast<const int4 *>(s_d + p);
                const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + o.x));
                const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(base + o.y));
                const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(base + o.z));
                const uint4 s3 = __ldg(reinterpret_cast<const uint4 *>(base + o.w));
                A0 |= (s0.x | s1.x) | (s2.x | s3.x); A1 |= (s0.y | s1.y) | (s2.y | s3.y);
                A2 |= (s0.z | s1.z) | (s2.z | s3.z); A3 |= (s0.w | s1.w) | (s2.w | s3.w);
            }
            for (; p < e2; p += 4) {                        // lower output only
                const int4 o = *reinterpret_cast<const int4 *>(s_d + p);
                const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + o.x));
                const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(base + o.y));
                const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(base + o.z));
                const uint4 s3 = __ldg(reinterpret_cast<const uint4 *>(base + o.w));
                B0 |= (s0.x | s1.x) | (s2.x | s3.x); B1 |= (s0.y | s1.y) | (s2.y | s3.y);
                B2 |= (s0.z | s1.z) | (s2.z | s3.z); B3 |= (s0.w | s1.w) | (s2.w | s3.w);
            }
        }
        const int64_t wi = ((int64_t)z * ny + y0) * NW + 4 * lr;
        bb_commit4(written, idx, wi, A0, A1, A2, A3, val);
        bb_commit4(written, idx, wi + NW, B0, B1, B2, B3, val);
        return;
    }
    // ---- border row groups: the one-output loop with bounds checks for each of the two rows
    for (int o = 0; o < 2; ++o) {
        const int y = y0 + o;
        if (y >= ny) continue;
        const uint32_t *base = seeds + ((int64_t)zs * ny + y) * NW + 4 * lr;
        uint32_t A0 = 0, A1 = 0, A2 = 0, A3 = 0;
        int p = 0;
        for (int a = bp.W; a >= 0; --a) {
            if (a < bp.W) BB_DILATE4(A0, A1, A2, A3)
            const int pend = bp.ring_end[a];
            for (; p < pend; ++p) {
                const int e = s_dyz[p];
                const int yy = y + (int)(short)(e & 0xFFFF), zz = zs + (e >> 16);
                if ((unsigned)yy < (unsigned)ny && (unsigned)zz < (unsigned)nz_src) {
                    const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(base + s_off[p]));
                    A0 |= s0.x; A1 |= s0.y; A2 |= s0.z; A3 |= s0.w;
                }
            }
        }
        const int64_t wi = ((int64_t)z * ny + y) * NW + 4 * lr;
        bb_commit4(written, idx, wi, A0, A1, A2, A3, val);
    }
#undef BB_DILATE4
}
