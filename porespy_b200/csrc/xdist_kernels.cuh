// xdist_kernels.cuh -- binary first pass of every separable transform on this path: distance
// along x (the contiguous axis) from each voxel to the nearest "site" of its own line.
//
//   XD_EDT : site <=> in[x] == 0            -> uint16 distance, >= 0x8000 when the line has no
//            site (first pass of edt.edt, /root/reference/src/porespy/filters/_funcs.py:1126)
//   XD_LT  : site <=> in[x] <= k (class map: seeds of radius k, F:1180/1196)
//                                            -> uint8 min(distance, cap), cap = W + 1 <= 254
//            (first pass of the per-radius dilation that replaces edt(~seeds) < r, F:1191)
//
// One warp per line, a lane owns 16-voxel chunks (one 16-byte load).  Sites are found with
// byte-SWAR compares; the chunk's 16 distances come from two 8-step running-distance
// recurrences evaluated two voxels at a time (voxel j and j+8 in the halves of one register)
// with the sm_100a DPX instruction VIADDMNMX.U16x2  (run = min(run + 1, site ? 0 : 0xFFFF)).
// Chunk-to-chunk carries (nearest site before / after the chunk) use warp-shuffle scans.
#pragma once
#include "common.cuh"

#define XD_EDT 0
#define XD_LT 1

// bit 7 of every byte of the result is set iff that byte of w is >= n  (1 <= n <= 254)
__device__ __forceinline__ uint32_t swar_ge(uint32_t w, uint32_t nl4, uint32_t sel)
{
    // t.bit7 = (low 7 bits of byte) >= (low 7 bits of n); no borrow crosses a byte
    const uint32_t t = ((w & 0x7F7F7F7Fu) | 0x80808080u) - nl4;
    // n < 128 (sel = ~0): byte >= n <=> bit7(w) | t ;  n >= 128 (sel = 0): bit7(w) & t
    return ((t & w) | ((t | w) & sel)) & 0x80808080u;
}

// prmt.b32 with the full 4-bit selectors (bit 3 of a selector nibble replicates the sign of the
// selected byte); __byte_perm only documents 3-bit selectors
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 4-bit mask of the bytes whose bit 7 is set
__device__ __forceinline__ uint32_t gather_bit7(uint32_t r)
{
    return (((r >> 7) * 0x00204081u) >> 21) & 0xFu;
}

template <int MODE>
__global__ void __launch_bounds__(256)
xdist_kernel(const uint8_t *__restrict__ in, void *__restrict__ outp, int64_t nlines, int nx,
             int k, int cap, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    extern __shared__ int xd_smem[];
    const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    const int nch = (nx + 15) >> 4;
    int *lastp = xd_smem + (size_t)wid * 2 * nch;     // last site at or before the end of chunk c
    int *firstp = lastp + nch;                        // first site at or after the start of chunk c
    const int NONE_L = -0x8000, NONE_R = 0x7FFF + nx; // "site" positions that give distance 0x7FFF at x=-1 / x=nx
    const uint32_t n = MODE == XD_EDT ? 1u : (uint32_t)(k + 1);   // non-site <=> byte >= n
    const uint32_t nl4 = (n & 0x7Fu) * 0x01010101u;
    const uint32_t sel = n < 128u ? 0xFFFFFFFFu : 0u;
    const bool vec = (nx & 15) == 0 && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(outp) & 15u) == 0);

    for (int64_t line = (int64_t)blockIdx.x * warps + wid; line < nlines; line += (int64_t)gridDim.x * warps) {
        const uint8_t *row = in + line * nx;
        // ---- phase 1: per-chunk site masks -> positions of the last / first site, scanned over chunks
        int carry = NONE_L;
        for (int base = 0; base < nch; base += 32) {
            const int c = base + lane;
            uint32_t mk = 0;
            if (c < nch) {
                uint4 v;
                if (vec) v = __ldg(reinterpret_cast<const uint4 *>(row) + c);
                else {
                    uint32_t w4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) w4[q] = load4(row, 16 * c + 4 * q, nx, 0xFFu);
                    v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
                const uint32_t s0 = swar_ge(v.x, nl4, sel) ^ 0x80808080u, s1 = swar_ge(v.y, nl4, sel) ^ 0x80808080u;
                const uint32_t s2 = swar_ge(v.z, nl4, sel) ^ 0x80808080u, s3 = swar_ge(v.w, nl4, sel) ^ 0x80808080u;
                mk = gather_bit7(s0) | (gather_bit7(s1) << 4) | (gather_bit7(s2) << 8) | (gather_bit7(s3) << 12);
                if (!vec && 16 * c + 16 > nx) mk &= (1u << (nx - 16 * c)) - 1u;     // padding is never a site
            }
            int vl = mk ? 16 * c + 31 - __clz(mk) : NONE_L;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, vl, off);
                if (lane >= off) vl = max(vl, t);
            }
            vl = max(vl, carry);
            carry = __shfl_sync(0xFFFFFFFFu, vl, 31);
            if (c < nch) {
                lastp[c] = vl;
                firstp[c] = mk ? 16 * c + __ffs(mk) - 1 : NONE_R;
            }
        }
        __syncwarp();
        carry = NONE_R;
        for (int base = ((nch - 1) / 32) * 32; base >= 0; base -= 32) {
            const int c = base + lane;
            int vr = (c < nch) ? firstp[c] : NONE_R;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_down_sync(0xFFFFFFFFu, vr, off);
                if (lane + off < 32) vr = min(vr, t);
            }
            vr = min(vr, carry);
            carry = __shfl_sync(0xFFFFFFFFu, vr, 0);
            if (c < nch) firstp[c] = vr;
        }
        __syncwarp();
        // ---- phase 2: the 16 distances of every chunk
        for (int c = lane; c < nch; c += 32) {
            uint4 v;
            if (vec) v = __ldg(reinterpret_cast<const uint4 *>(row) + c);      // L1/L2 hit
            else {
                uint32_t w4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) w4[q] = load4(row, 16 * c + 4 * q, nx, 0xFFu);
                v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            uint32_t ns[4] = {swar_ge(v.x, nl4, sel), swar_ge(v.y, nl4, sel), swar_ge(v.z, nl4, sel),
                              swar_ge(v.w, nl4, sel)};                         // bit7 set <=> NOT a site
            if (!vec && 16 * c + 16 > nx) {                                    // padding: non-site
                const int valid = nx - 16 * c;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (4 * q + b >= valid) ns[q] |= 0x80u << (8 * b);
            }
            const uint32_t mlo = (gather_bit7(ns[0] ^ 0x80808080u) | (gather_bit7(ns[1] ^ 0x80808080u) << 4));
            const uint32_t mhi = (gather_bit7(ns[2] ^ 0x80808080u) | (gather_bit7(ns[3] ^ 0x80808080u) << 4));
            const int Lpos = c > 0 ? lastp[c - 1] : NONE_L;
            const int Rpos = c + 1 < nch ? firstp[c + 1] : NONE_R;
            const uint32_t carryL = (uint32_t)min(16 * c - 1 - Lpos, 0x7FFF + 16 * c);     // distance at x = 16c - 1
            const uint32_t carryR = (uint32_t)min(Rpos - (16 * c + 16), 0x7FFF + nx - 16 * c - 16);  // at x = 16c + 16
            // halves: low 16 bits follow voxel j, high 16 bits voxel j + 8
            const uint32_t fhi = mlo ? (uint32_t)(7 - (31 - __clz(mlo))) : carryL + 8u;    // forward value at voxel 7
            const uint32_t blo = mhi ? (uint32_t)(__ffs(mhi) - 1) : carryR + 8u;           // backward value at voxel 8
            uint32_t s2[8], f2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = j & 3;
                const uint32_t selp = (8u | p) | ((8u | p) << 4) | ((12u | p) << 8) | ((12u | p) << 12);
                s2[j] = prmt(ns[j >> 2], ns[(j >> 2) + 2], selp);       // 0xFFFF non-site, 0 site
            }
            uint32_t run = carryL | (fhi << 16);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                run = __viaddmin_u16x2(run, 0x00010001u, s2[j]);
                f2[j] = run;
            }
            run = blo | (carryR << 16);
            uint32_t d2[8];
            const uint32_t cap2 = (uint32_t)cap * 0x00010001u;
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                run = __viaddmin_u16x2(run, 0x00010001u, s2[j]);
                d2[j] = __vminu2(f2[j], run);
                if (MODE == XD_LT) d2[j] = __vminu2(d2[j], cap2);
            }
            if (MODE == XD_EDT) {
                uint16_t *orow = reinterpret_cast<uint16_t *>(outp) + line * nx;
                uint4 o0, o1;
                o0.x = __byte_perm(d2[0], d2[1], 0x5410); o0.y = __byte_perm(d2[2], d2[3], 0x5410);
                o0.z = __byte_perm(d2[4], d2[5], 0x5410); o0.w = __byte_perm(d2[6], d2[7], 0x5410);
                o1.x = __byte_perm(d2[0], d2[1], 0x7632); o1.y = __byte_perm(d2[2], d2[3], 0x7632);
                o1.z = __byte_perm(d2[4], d2[5], 0x7632); o1.w = __byte_perm(d2[6], d2[7], 0x7632);
                if (vec) {
                    reinterpret_cast<uint4 *>(orow)[2 * c] = o0;
                    reinterpret_cast<uint4 *>(orow)[2 * c + 1] = o1;
                } else {
                    const uint32_t w8[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (16 * c + j < nx) orow[16 * c + j] = (uint16_t)(w8[j >> 1] >> (16 * (j & 1)));
                }
            } else {
                uint8_t *orow = reinterpret_cast<uint8_t *>(outp) + line * nx;
                const uint32_t t01 = __byte_perm(d2[0], d2[1], 0x6240), t23 = __byte_perm(d2[2], d2[3], 0x6240);
                const uint32_t t45 = __byte_perm(d2[4], d2[5], 0x6240), t67 = __byte_perm(d2[6], d2[7], 0x6240);
                uint4 o;
                o.x = __byte_perm(t01, t23, 0x5410); o.y = __byte_perm(t45, t67, 0x5410);
                o.z = __byte_perm(t01, t23, 0x7632); o.w = __byte_perm(t45, t67, 0x7632);
                if (vec) reinterpret_cast<uint4 *>(orow)[c] = o;
                else {
                    const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) store4(orow, 16 * c + 4 * q, nx, w4[q]);
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ x pass from seed BITS
// The per-radius x pass when the seed set of the radius is already packed (lt_packn_kernel packs up to 16 radii
// from one read of the class map): a lane owns one 32-voxel word, so the site masks come for free and the
// byte-SWAR compares -- most of xdist_kernel<XD_LT>'s instructions -- disappear; the distances are the same two
// running-distance recurrences (VIADDMNMX.U16x2, voxel j and j + 16 in the halves of one register).
//   gx[line][x] = min(distance along x to the nearest set bit of the line, cap)      cap = W + 1 <= 254
//   xflag[line][w] (optional) = 1 iff some voxel of word w has gx < cap
// nx = 32 * nw.  dyn smem = warps * 2 * nw ints.
__global__ void __launch_bounds__(256)
xdist_bits_kernel(const uint32_t *__restrict__ bits, uint8_t *__restrict__ gx, int64_t nlines, int nw, int cap,
                  const int *__restrict__ gate, uint8_t *__restrict__ xflag)
{
    if (gate && *gate == 0) return;
    extern __shared__ int xb_smem[];
    const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    int *lastp = xb_smem + (size_t)wid * 2 * nw;      // last site at or before the end of word w
    int *firstp = lastp + nw;                         // first site at or after the start of word w
    const int nx = 32 * nw;
    const int NONE_L = -0x4000, NONE_R = 0x4000 + nx;
    const uint32_t cap2 = (uint32_t)cap * 0x00010001u;
    for (int64_t line = (int64_t)blockIdx.x * warps + wid; line < nlines; line += (int64_t)gridDim.x * warps) {
        const uint32_t *row = bits + line * nw;
        int carry = NONE_L;
        for (int base = 0; base < nw; base += 32) {
            const int w = base + lane;
            const uint32_t m = w < nw ? __ldg(row + w) : 0u;
            int vl = m ? 32 * w + 31 - __clz(m) : NONE_L;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, vl, off);
                if (lane >= off) vl = max(vl, t);
            }
            vl = max(vl, carry);
            carry = __shfl_sync(0xFFFFFFFFu, vl, 31);
            if (w < nw) {
                lastp[w] = vl;
                firstp[w] = m ? 32 * w + __ffs(m) - 1 : NONE_R;
            }
        }
        __syncwarp();
        carry = NONE_R;
        for (int base = ((nw - 1) / 32) * 32; base >= 0; base -= 32) {
            const int w = base + lane;
            int vr = w < nw ? firstp[w] : NONE_R;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_down_sync(0xFFFFFFFFu, vr, off);
                if (lane + off < 32) vr = min(vr, t);
            }
            vr = min(vr, carry);
            carry = __shfl_sync(0xFFFFFFFFu, vr, 0);
            if (w < nw) firstp[w] = vr;
        }
        __syncwarp();
        for (int w = lane; w < nw; w += 32) {
            const uint32_t m = __ldg(row + w);
            const int Lpos = w > 0 ? lastp[w - 1] : NONE_L;
            const int Rpos = w + 1 < nw ? firstp[w + 1] : NONE_R;
            const uint32_t carryL = (uint32_t)min(32 * w - 1 - Lpos, 0x4000);       // distance at x = 32w - 1
            const uint32_t carryR = (uint32_t)min(Rpos - (32 * w + 32), 0x4000);    // distance at x = 32w + 32
            const uint32_t mlo = m & 0xFFFFu, mhi = m >> 16;
            const uint32_t fhi = mlo ? (uint32_t)(15 - (31 - __clz(mlo))) : carryL + 16u;   // forward value at voxel 15
            const uint32_t blo = mhi ? (uint32_t)(__ffs(mhi) - 1) : carryR + 16u;           // backward value at voxel 16
            uint32_t f[16];
            uint32_t run = carryL | (fhi << 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t s = (~(m >> j) & 0x00010001u) * 0xFFFFu;             // 0xFFFF per non-site half
                run = __viaddmin_u16x2(run, 0x00010001u, s);
                f[j] = run;
            }
            run = blo | (carryR << 16);
            uint32_t d[16];
#pragma unroll
            for (int j = 15; j >= 0; --j) {
                const uint32_t s = (~(m >> j) & 0x00010001u) * 0xFFFFu;
                run = __viaddmin_u16x2(run, 0x00010001u, s);
                d[j] = __vminu2(__vminu2(f[j], run), cap2);
            }
            if (xflag) {
                // activity flag of the word: some voxel lies within W = cap - 1 of a seed of its line (the y pass
                // skips tiles and rows without any)
                uint32_t mm = d[0];
#pragma unroll
                for (int j = 1; j < 16; ++j) mm = __vminu2(mm, d[j]);
                xflag[line * nw + w] = min(mm & 0xFFFFu, mm >> 16) < (uint32_t)cap ? 1 : 0;
            }
            uint32_t olo[4], ohi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t t01 = __byte_perm(d[4 * q], d[4 * q + 1], 0x6240), t23 = __byte_perm(d[4 * q + 2], d[4 * q + 3], 0x6240);
                olo[q] = __byte_perm(t01, t23, 0x5410);          // voxels 4q .. 4q + 3
                ohi[q] = __byte_perm(t01, t23, 0x7632);          // voxels 16 + 4q .. 16 + 4q + 3
            }
            uint4 *o = reinterpret_cast<uint4 *>(gx + line * nx + 32 * w);
            o[0] = make_uint4(olo[0], olo[1], olo[2], olo[3]);
            o[1] = make_uint4(ohi[0], ohi[1], ohi[2], ohi[3]);
        }
        __syncwarp();
    }
}
