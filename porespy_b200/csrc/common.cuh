// common.cuh -- shared helpers for the psb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PSB_INF 0xFFFFFFFFu

// Class-map byte values (see lt_kernels.cuh)
#define CLS_BG 255u        // background voxel (d2 == 0): never a seed, never filled
#define CLS_NEVER 254u     // foreground voxel that is a seed for no threshold of this call
// x-distance byte values inside the xy tile
#define GX_BG 255u         // background voxel: output skipped
#define GX_FAR 254u        // no seed within 253 voxels along x

#include <vector>
struct ProfRec {
    int kid;
    cudaEvent_t a, b;
};

struct psb200_ctx {
    int device;
    int sm_count;
    int max_smem_optin;
    int algo;
    int profile;
    int bit_tmax;      // thresholds T <= bit_tmax use the bit-parallel dilation (0: never)
    int bit4;          // bit path: four-words-per-lane kernel for rows of 32 / 64 / 128 words
    int foot;          // warp footprint of the 16-bit EDT min-plus scans: 0 = 64 x 8 voxels, 1 = 32 x 16 (default: 3 % faster, r2b)
    int uf_records;    // flood: row-rooted forest + link records (default) / per-voxel job lists
    int yflags;        // byte path: the bit-based x pass leaves an activity byte per word, the y pass skips idle tiles / rows
    int zwide;         // z sweeps with 8 columns per thread and 16 planes in flight (default) / 4 columns, 8 planes
    int edt_h;         // halo rows staged on each side of a 128-row tile of the 16-bit EDT passes (scans beyond it read global
                       // memory); 32: 52 KB of shared memory, 4 resident blocks per SM -- y 3.48 -> 3.20, z 3.18 -> 2.71 ms vs 48 (r2m)
    int ydirect;       // per-radius y pass: reach bytes straight from registers to global memory (default) / via a tile copy
    int bitquad;       // bit path: two output rows per lane (lt_bitball4d_kernel, default) / one row per lane
    int xbits;         // per-radius x pass from packed seed bits (xdist_bits_kernel, default) or from the class map
    int ycoarse;       // per-radius y pass: hierarchical scan (lt_y3_kernel, default) or the plain one (lt_y2_kernel)
    int edt16;         // EDT y/z passes: 16-bit two-voxels-per-instruction kernel first, uint32 kernel gated behind it
    int *flags;        // 64 device ints owned by the context (overflow flags of the 16-bit passes)
    unsigned flag_slot;
    long long launches;
    std::vector<ProfRec> prof;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t byte_of(uint32_t v, int j) { return (v >> (8 * j)) & 0xFFu; }

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return a | (b << 8) | (c << 16) | (d << 24);
}

// Load 4 consecutive bytes row[x..x+3] of a line of length nx; bytes outside [0,nx) read as
// `fill`.  Uses one 32-bit load when the address is 4-byte aligned and fully inside.
__device__ __forceinline__ uint32_t load4(const uint8_t *__restrict__ row, int x, int nx,
                                          uint32_t fill)
{
    if (x >= 0 && x + 3 < nx && ((reinterpret_cast<uintptr_t>(row + x) & 3u) == 0))
        return __ldg(reinterpret_cast<const uint32_t *>(row + x));
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int xx = x + j;
        uint32_t b = (xx >= 0 && xx < nx) ? (uint32_t)__ldg(row + xx) : fill;
        v |= b << (8 * j);
    }
    return v;
}

__device__ __forceinline__ void store4(uint8_t *__restrict__ row, int x, int nx, uint32_t v)
{
    if (x + 3 < nx && ((reinterpret_cast<uintptr_t>(row + x) & 3u) == 0)) {
        *reinterpret_cast<uint32_t *>(row + x) = v;
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (x + j < nx) row[x + j] = (uint8_t)byte_of(v, j);
}

// ceil(sqrt(x)) for 1 <= x <= 65536 (exact: sqrtf is correctly rounded and the gap between
// sqrt of a non-square and the next integer is > 1/(2*257) >> ulp).
__device__ __forceinline__ uint32_t ceil_sqrt_small(uint32_t x)
{
    return (uint32_t)__float2int_ru(sqrtf((float)x));
}
