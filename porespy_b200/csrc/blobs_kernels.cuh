// blobs_kernels.cuh -- device-side `ps.generators.blobs`
// (/root/reference/src/porespy/generators/_imgen.py:1023-1051, norm_to_uniform
// /root/reference/src/porespy/tools/_funcs.py:963-969): the input generator of every benchmark
// configuration.  Host blobs() needs >= 4 float64 temporaries of the volume
// (69 GB each at 2048^3) and minutes of single-threaded scipy; here the field never leaves HBM.
//
//   noise (float64 uniform)  ->  gaussian_filter (three 1-D correlations, axis 0, 1, 2 like
//   scipy.ndimage.gaussian_filter, mode='reflect', truncate=4)  ->  (f - mean) / std  ->
//   0.5 erfc(-z / sqrt 2)  ->  (c - min) / (max - min)  ->  `< porosity`
//
// Arithmetic is float64 in scipy's own order (ni_filters.c NI_Correlate1D, symmetric branch:
// centre product first, then (a[-j] + a[j]) * w[j] from the outermost tap inwards, no fused
// multiply-add), so for the same noise the filtered field is bit-equal to scipy's; mean / std are
// fixed-order per-plane sums (numpy's pairwise order is not reproduced), so the thresholded image can
// differ from the host blobs() only where the uniformised value is within an ulp of `porosity`.
//
// Noise: either uploaded by the caller (numpy's seeded MT19937 stream: the same image as the
// reference for the same seed) or Philox4x32-10 keyed by (seed, GLOBAL voxel index): a pure function of
// the voxel, so z-slab shards generate their own planes (plus the filter halo) without communication
// and the global image does not depend on the number of GPUs.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// 53-bit uniform in [0, 1) from two 32-bit words (the construction numpy's random_sample uses)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// out[i] = U(seed, first + i), i in [0, n): element e of the global index space comes from Philox counter
// e >> 1, half e & 1
__global__ void __launch_bounds__(256)
noise_philox_kernel(double *__restrict__ out, int64_t n, uint64_t seed, uint64_t first)
{
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t p0 = first >> 1, p1 = (first + (uint64_t)n + 1) >> 1;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += step) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)p, (uint32_t)(p >> 32), 0u, 0u), key);
        const uint64_t e = 2 * p;
        if (e >= first && e < first + (uint64_t)n) out[e - first] = u53(r.x, r.y);
        if (e + 1 >= first && e + 1 < first + (uint64_t)n) out[e + 1 - first] = u53(r.z, r.w);
    }
}

// ------------------------------------------------------------------------- 1-D correlations
// index of the half-sample symmetric extension  (d c b a | a b c d | d c b a)  of a line of n samples
__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t n)
{
    const int64_t per = 2 * n;
    int64_t m = i % per;
    if (m < 0) m += per;
    return m >= n ? per - 1 - m : m;
}

// scipy's symmetric correlation at one sample: `at(j)` returns the extended line at offset j
template <typename At>
__device__ __forceinline__ double corr_sym(const double *__restrict__ w, int radius, At at)
{
    double acc = __dmul_rn(at(0), w[radius]);
    for (int j = radius; j >= 1; --j)
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(-j), at(j)), w[radius - j]));
    return acc;
}

// axis 2 (contiguous): one block per (line, segment of GX_SEG samples); the segment plus its halo is staged
// in shared memory.  dyn smem = (GX_SEG + 2 radius + radius + 1) doubles.
#define GX_SEG 1024
__global__ void __launch_bounds__(256)
gauss_x_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t nlines, int nx, int radius,
               const double *__restrict__ wg)
{
    extern __shared__ double gsm[];
    double *w = gsm, *seg = gsm + radius + 1;
    for (int i = threadIdx.x; i <= radius; i += blockDim.x) w[i] = wg[i];
    const int nseg = (nx + GX_SEG - 1) / GX_SEG;
    for (int64_t job = blockIdx.x; job < nlines * nseg; job += gridDim.x) {
        const int64_t line = job / nseg;
        const int x0 = (int)(job % nseg) * GX_SEG;
        const int len = min(GX_SEG, nx - x0);
        const double *row = in + line * nx;
        __syncthreads();
        for (int i = threadIdx.x; i < len + 2 * radius; i += blockDim.x)
            seg[i] = row[reflect_index((int64_t)x0 - radius + i, nx)];
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const double *c = seg + radius + i;
            out[line * nx + x0 + i] = corr_sym(w, radius, [&](int j) { return c[j]; });
        }
    }
}

// axis 1 / axis 0 (strided): a block owns 32 adjacent columns x GC_ROWS outputs along the axis; the
// (GC_ROWS + 2 radius) x 32 tile is staged in shared memory.
//   ncols    : number of independent columns inside one outer slice (nx for axis 1, ny*nx for axis 0)
//   stride   : distance between consecutive samples along the axis
//   n_out    : outputs along the axis, sample o of the output is global sample g0_out + o
//   n_in     : input samples along the axis; input sample i is global sample g0_in + i
//   n_glob   : length of the global line (reflection happens at its ends)
//   outer, ostride_in / ostride_out : outer slices (nz for axis 1, 1 for axis 0)
__global__ void __launch_bounds__(256)
gauss_col_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t ncols, int64_t stride,
                 int n_out, int64_t g0_out, int n_in, int64_t g0_in, int64_t n_glob, int64_t outer,
                 int64_t ostride_in, int64_t ostride_out, int radius, int rows, const double *__restrict__ wg)
{
    extern __shared__ double gsm[];
    double *w = gsm, *tile = gsm + radius + 1;          // tile[(rows + 2 radius)][32]
    for (int i = threadIdx.x; i <= radius; i += blockDim.x) w[i] = wg[i];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t ctiles = (ncols + 31) / 32;
    const int64_t rtiles = (n_out + rows - 1) / rows;
    const int64_t njobs = ctiles * rtiles * outer;
    for (int64_t job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int64_t rt = job % rtiles, ct = (job / rtiles) % ctiles, o = job / (rtiles * ctiles);
        const int64_t col = ct * 32 + lane;
        const int r0 = (int)rt * rows;
        const int nr = min(rows, n_out - r0);
        __syncthreads();
        for (int r = wrp; r < nr + 2 * radius; r += nwarps) {
            const int64_t g = reflect_index(g0_out + r0 - radius + r, n_glob);     // global sample
            double v = 0.0;
            if (col < ncols) v = in[o * ostride_in + (g - g0_in) * stride + col];
            tile[r * 32 + lane] = v;
        }
        __syncthreads();
        if (col < ncols)
            for (int r = wrp; r < nr; r += nwarps) {
                const double *c = tile + (size_t)(r + radius) * 32 + lane;
                out[o * ostride_out + (int64_t)(r0 + r) * stride + col] =
                    corr_sym(w, radius, [&](int j) { return c[j * 32]; });
            }
    }
    (void)n_in;
}

// --------------------------------------------------------------------------------- statistics
// Fixed-order partial reductions: part[plane][chunk] over ST_CHUNKS equal chunks of every plane, so the
// host sums the same numbers in the same order whatever the sharding.
//   mode 0: sum(x);  mode 1: sum((x - mean)^2);  mode 2: min(x);  mode 3: max(x)
#define ST_CHUNKS 16
__global__ void __launch_bounds__(256)
stats_kernel(const double *__restrict__ x, int64_t nplanes, int64_t plane, double mean, int mode,
             double *__restrict__ part)
{
    __shared__ double sh[8];
    for (int64_t job = blockIdx.x; job < nplanes * ST_CHUNKS; job += gridDim.x) {
        const int64_t p = job / ST_CHUNKS, c = job % ST_CHUNKS;
        const int64_t per = (plane + ST_CHUNKS - 1) / ST_CHUNKS;
        const int64_t i0 = c * per, i1 = min(plane, i0 + per);
        double acc = mode == 2 ? INFINITY : (mode == 3 ? -INFINITY : 0.0);
        for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
            const double v = x[p * plane + i];
            if (mode == 0) acc = __dadd_rn(acc, v);
            else if (mode == 1) { const double d = __dadd_rn(v, -mean); acc = __dadd_rn(acc, __dmul_rn(d, d)); }
            else if (mode == 2) acc = fmin(acc, v);
            else acc = fmax(acc, v);
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double o = __shfl_xor_sync(0xFFFFFFFFu, acc, s);
            acc = mode == 2 ? fmin(acc, o) : (mode == 3 ? fmax(acc, o) : __dadd_rn(acc, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = sh[0];
            for (int i = 1; i < 8; ++i) t = mode == 2 ? fmin(t, sh[i]) : (mode == 3 ? fmax(t, sh[i]) : __dadd_rn(t, sh[i]));
            part[job] = t;
        }
    }
}

// ------------------------------------------------------------------------ norm_to_uniform + threshold
__device__ __forceinline__ double uniformise(double f, double mean, double sd)
{
    const double z = __ddiv_rn(__dadd_rn(f, -mean), sd);                  // (im - mean) / std
    return __dmul_rn(0.5, erfc(__ddiv_rn(-z, 1.4142135623730951)));      // 1/2 * erfc(-im / sqrt(2))
}

// porosity > 0: out8[i] = u < porosity;  else outf[i] = u   (u = (c - cmin) / (cmax - cmin) * 1 + 0)
__global__ void __launch_bounds__(256)
blobs_finish_kernel(const double *__restrict__ f, int64_t n, double mean, double sd, double fmin_, double fmax_,
                    double porosity, uint8_t *__restrict__ out8, double *__restrict__ outf)
{
    const double cmin = uniformise(fmin_, mean, sd), cmax = uniformise(fmax_, mean, sd);
    const double span = __dadd_rn(cmax, -cmin);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double c = uniformise(f[i], mean, sd);
        double u = __ddiv_rn(__dadd_rn(c, -cmin), span);
        u = __dadd_rn(__dmul_rn(u, 1.0), 0.0);
        if (out8) out8[i] = u < porosity ? 1 : 0;
        else outf[i] = u;
    }
}
