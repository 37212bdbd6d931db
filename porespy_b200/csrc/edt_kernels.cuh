// edt_kernels.cuh -- exact squared Euclidean distance transform, separable passes.
//
// Replaces the third-party `edt.edt` PoreSpy calls at
// /root/reference/src/porespy/filters/_funcs.py:1126 and :1191 (black_border=False).
//
//   pass x : one warp per line.  The line's "site" voxels (zeros of the image, or seeds of
//            the current radius) are turned into a bit mask with warp ballots, a warp
//            shuffle scan gives every 32-voxel word the nearest site before/after it, and
//            every voxel gets its distance with clz/ffs on its own word.  Coalesced.
//   pass y/z: one thread per line, lanes along x so every global access is a coalesced
//            row segment.  Lower envelope of parabolas in Meijster's all-integer form; the
//            envelope stack lives in a global scratch buffer laid out [depth][thread] so
//            pushes/pops of neighbouring lanes coalesce, with the stack top cached in
//            registers (the common no-pop step touches no scratch memory).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------ pass x
// SITE_MODE 0: site <=> in[x] == 0           (EDT of the image itself)
// SITE_MODE 1: site <=> in[x] <= k           (EDT of ~seeds, seeds = class <= k)
template <int SITE_MODE>
__device__ __forceinline__ bool is_site(uint32_t v, int k)
{
    return SITE_MODE == 0 ? (v == 0u) : (v <= (uint32_t)k);
}

#define XS_NONE_R 0x3FFFFFFF

template <int SITE_MODE>
__global__ void __launch_bounds__(256)
edt_x_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, int64_t nlines, int nx,
             int k)
{
    extern __shared__ uint32_t xs_smem[];
    const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    const int nwords = (nx + 31) >> 5;
    uint32_t *words = xs_smem + (size_t)wid * 3 * nwords;
    int *lastz = reinterpret_cast<int *>(words + nwords);   // last site at or before end of word w
    int *nextz = lastz + nwords;                            // first site at or after start of word w

    for (int64_t line = (int64_t)blockIdx.x * warps + wid; line < nlines;
         line += (int64_t)gridDim.x * warps) {
        const uint8_t *p = in + line * nx;
        for (int c = 0; c < nwords; ++c) {
            int x = c * 32 + lane;
            bool s = (x < nx) && is_site<SITE_MODE>(p[x], k);
            uint32_t m = __ballot_sync(0xFFFFFFFFu, s);
            if (lane == 0) words[c] = m;
        }
        __syncwarp();
        // inclusive max-scan of "position of last site" over words
        int carry = -1;
        for (int base = 0; base < nwords; base += 32) {
            int w = base + lane;
            uint32_t wd = (w < nwords) ? words[w] : 0u;
            int v = wd ? (32 * w + 31 - __clz(wd)) : -1;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_up_sync(0xFFFFFFFFu, v, off);
                if (lane >= off) v = max(v, t);
            }
            v = max(v, carry);
            if (w < nwords) lastz[w] = v;
            carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        }
        // inclusive (reverse) min-scan of "position of first site"
        carry = XS_NONE_R;
        for (int base = ((nwords - 1) / 32) * 32; base >= 0; base -= 32) {
            int w = base + lane;
            uint32_t wd = (w < nwords) ? words[w] : 0u;
            int v = wd ? (32 * w + __ffs(wd) - 1) : XS_NONE_R;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_down_sync(0xFFFFFFFFu, v, off);
                if (lane + off < 32) v = min(v, t);
            }
            v = min(v, carry);
            if (w < nwords) nextz[w] = v;
            carry = __shfl_sync(0xFFFFFFFFu, v, 0);
        }
        __syncwarp();
        uint32_t *o = out + line * nx;
        for (int x = lane; x < nx; x += 32) {
            int w = x >> 5, b = x & 31;
            uint32_t wd = words[w];
            uint32_t ml = wd & (0xFFFFFFFFu >> (31 - b));
            int L = ml ? (32 * w + 31 - __clz(ml)) : (w > 0 ? lastz[w - 1] : -1);
            uint32_t mr = wd & (0xFFFFFFFFu << b);
            int R = mr ? (32 * w + __ffs(mr) - 1) : (w + 1 < nwords ? nextz[w + 1] : XS_NONE_R);
            uint32_t d = 0x7FFFFFFFu;
            if (L >= 0) d = (uint32_t)(x - L);
            if (R != XS_NONE_R) d = min(d, (uint32_t)(R - x));
            o[x] = (d == 0x7FFFFFFFu) ? PSB_INF : d * d;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------- pass y/z
struct EdtStoreU32 {
    uint32_t *out;
    __device__ __forceinline__ void operator()(int64_t i, uint32_t d2) const { out[i] = d2; }
};

// z-pass epilogue of the generic per-radius path: fill <=> d2 < T; write radius index once.
struct EdtStoreFill {
    uint8_t *idx;
    uint32_t T;
    uint8_t val;
    __device__ __forceinline__ void operator()(int64_t i, uint32_t d2) const
    {
        if (d2 < T && idx[i] == 0) idx[i] = val;
    }
};

// Column c of a pass: element u lives at in[(c / inner) * outer + (c % inner) + u * stride].
//   y pass: inner = nx,     outer = ny*nx, stride = nx,    n = ny, ncols = nz*nx
//   z pass: inner = ny*nx,  outer = 0,     stride = ny*nx, n = nz, ncols = ny*nx
// stk: uint2 scratch [n][nthreads]; entry = {g(apex), apex | (takeover << 16)}.
template <typename Store>
__global__ void __launch_bounds__(128)
edt_col_kernel(const uint32_t *in, Store store, int64_t ncols, int64_t inner,
               int64_t outer, int64_t stride, int n, uint2 *stk, const int *gate = nullptr)   // in may alias store.out (in place)
{
    if (gate && *gate == 0) return;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint2 *mystk = stk + tid;
    for (int64_t c = tid; c < ncols; c += nthreads) {
        const int64_t base = (c / inner) * outer + (c % inner);
        int q = -1;                // stack depth - 1
        uint32_t tg = 0;           // g at the apex of the top parabola
        int ts = 0, tt = 0;        // apex position, first abscissa where the top parabola wins
        for (int u0 = 0; u0 < n; u0 += 4) {
            uint32_t gbuf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                gbuf[j] = (u0 + j < n) ? in[base + (int64_t)(u0 + j) * stride] : PSB_INF;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t g = gbuf[j];
                const int u = u0 + j;
                if (g == PSB_INF) continue;     // infinite parabola: never on the envelope
                while (q >= 0) {
                    const int64_t a = tt - ts, b = tt - u;
                    if (a * a + (int64_t)tg > b * b + (int64_t)g) {
                        --q;
                        if (q >= 0) {
                            uint2 e = mystk[(int64_t)q * nthreads];
                            tg = e.x; ts = (int)(e.y & 0xFFFFu); tt = (int)(e.y >> 16);
                        }
                    } else break;
                }
                if (q < 0) {
                    q = 0; tg = g; ts = u; tt = 0;
                    mystk[0] = make_uint2(g, (uint32_t)u);
                } else {
                    // first integer abscissa where parabola u is strictly below parabola ts
                    const int64_t num = (int64_t)(u - ts) * (u + ts) + (int64_t)g - (int64_t)tg;
                    const int64_t den = 2 * (int64_t)(u - ts);
                    int64_t sep;
                    if (num >= 0 && num < 0x7FFFFFFFLL) sep = (int64_t)((uint32_t)num / (uint32_t)den);
                    else if (num >= 0) sep = num / den;
                    else sep = -((-num + den - 1) / den);
                    const int64_t w = sep + 1;
                    if (w < n) {
                        ++q; tg = g; ts = u; tt = (int)w;
                        mystk[(int64_t)q * nthreads] = make_uint2(g, (uint32_t)u | ((uint32_t)w << 16));
                    }
                }
            }
        }
        if (q < 0) {
            for (int u = 0; u < n; ++u) store(base + (int64_t)u * stride, PSB_INF);
        } else {
            for (int u = n - 1; u >= 0; --u) {
                const int d = u - ts;
                store(base + (int64_t)u * stride, (uint32_t)(d * d) + tg);
                if (u == tt && q > 0) {
                    --q;
                    uint2 e = mystk[(int64_t)q * nthreads];
                    tg = e.x; ts = (int)(e.y & 0xFFFFu); tt = (int)(e.y >> 16);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- pointwise
__global__ void sqrt_f32_kernel(const uint32_t *d2, float *out, int64_t n, const int *gate = nullptr)   // may run in place
{
    if (gate && *gate == 0) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        uint32_t v = d2[i];
        out[i] = (v == PSB_INF) ? __int_as_float(0x7F800000) : sqrtf((float)v);
    }
}

__global__ void max_u32_kernel(const uint32_t *__restrict__ d2, int64_t n, uint32_t *__restrict__ out,
                               const int *gate = nullptr)
{
    if (gate && *gate == 0) return;
    uint32_t m = 0;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        m = max(m, d2[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if (lane_id() == 0 && m) atomicMax(out, m);
}

// 1-D images: the x pass already holds the final distances; apply the fill rule pointwise.
__global__ void __launch_bounds__(256)
fill_from_d2_kernel(const uint32_t *__restrict__ d2, EdtStoreFill store, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) store(i, d2[i]);
}


// ---- pieces of the gated envelope fallback of the fast EDT (psb200.cu edt_fast): uint16 x-distances ->
// uint32 squared distances (PSB_INF: no site in the line), and a gated copy
__global__ void __launch_bounds__(256)
sq16_to_u32_kernel(const uint16_t *__restrict__ dx, uint32_t *__restrict__ out, int64_t n, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const uint32_t d = dx[i];
        out[i] = d >= 0x8000u ? PSB_INF : d * d;
    }
}

__global__ void __launch_bounds__(256)
copy_u32_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int64_t n, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dst[i] = src[i];
}
