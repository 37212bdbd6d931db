// drainage_kernels.cuh -- image-based drainage with gravity:
// /root/reference/src/porespy/simulations/_drainage.py:104-154 (the pressure loop) and the sphere painter
// /root/reference/src/porespy/tools/_sphere_insertions.py:327-385 (`_insert_disks_at_points`).
//
// Per applied pressure p the reference thresholds fn = pc + rho g h at p, keeps what is connected to the
// inlets (trim_disconnected_blobs: the flood kernels), and inserts at every NEWLY invaded voxel s a sphere of
// radius int(dt[s]) holding the value p wherever nothing was written yet.  The sphere of integer radius r
// (`_make_ball(r, smooth=True)`: sqrt(|o|^2) <= r - 0.001) is the digital ball {o : |o|^2 < r^2}, empty for
// r = 0.  A voxel v is therefore painted at this step iff
//        min over new seeds s of  |v - s|^2 - r(s)^2   <   0
// which is a separable lower envelope of parabolas with different heights (a power diagram): one bounded
// min-plus pass per axis on  f(s) = C - r(s)^2  (C = r_max^2 keeps the values non-negative; voxels that are not
// seeds enter with the value C, which can never produce a result below C, so every scan ends after at most
// r_max steps).  The y / z passes are the min-plus kernels of the EDT (minplus_kernels.cuh), the x pass is
// below.
#pragma once
#include "common.cuh"

// fn = pc + rgh in the reference's own precisions (F = simulations/_drainage.py):
//   pc  = c0 / (dt * voxel_size)                      F:107   float32 product (float64 when voxel_size is a
//                                                             float64 numpy scalar), float64 quotient; 0 outside im
//   h   = (edt(h0) + 1) * voxel_size                  F:113   edt(h0) = index along the first image axis
//   rgh = delta_rho * g * h                           F:114
struct DrainFn {
    double c0;          // -(ndim - 1) * sigma * cos(theta)
    double vs64;        // voxel_size as float64
    float vs32;         // ... as float32
    double rg64;        // delta_rho * g
    float rg32;
    int den64;          // dt * voxel_size evaluated in float64
    int h64;            // h evaluated in float64
    int rgh64;          // rgh evaluated in float64
    int use_pc;         // pc given by the caller (float64 array)
    int64_t inner;      // voxels per step of the first image axis
};

__device__ __forceinline__ double drain_fn_at(const DrainFn &q, float dt, bool pore, const double *__restrict__ pc_user,
                                              int64_t v)
{
    double pc;
    if (q.use_pc) pc = pc_user[v];
    else {
        const double den = q.den64 ? __dmul_rn((double)dt, q.vs64) : (double)__fmul_rn(dt, q.vs32);
        pc = __ddiv_rn(q.c0, den);
    }
    if (!pore) pc = 0.0;                                               // F:108
    const float zf = (float)(v / q.inner);
    const double h = q.h64 ? __dmul_rn((double)__fadd_rn(zf, 1.0f), q.vs64) : (double)__fmul_rn(__fadd_rn(zf, 1.0f), q.vs32);
    const double rgh = q.rgh64 ? __dmul_rn(q.rg64, h) : (double)__fmul_rn(q.rg32, (float)h);
    return __dadd_rn(pc, rgh);                                         // F:115
}

// partial[2 b] = max{fn : fn < inf}, partial[2 b + 1] = min{fn over pore voxels : fn > -inf}   (F:122-123)
__global__ void __launch_bounds__(256)
drain_stats_kernel(const float *__restrict__ dt, const uint8_t *__restrict__ im, const double *__restrict__ pc_user,
                   int64_t n, const __grid_constant__ DrainFn q, double *__restrict__ partial)
{
    __shared__ double shmax[8], shmin[8];
    double vmax = -INFINITY, vmin = INFINITY;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const bool pore = im[v] != 0;
        const double f = drain_fn_at(q, dt[v], pore, pc_user, v);
        if (f < INFINITY) vmax = fmax(vmax, f);                        // (NaN compares false: skipped, like numpy's mask)
        if (pore && f > -INFINITY) vmin = fmin(vmin, f);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        vmax = fmax(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, s));
        vmin = fmin(vmin, __shfl_xor_sync(0xFFFFFFFFu, vmin, s));
    }
    if ((threadIdx.x & 31) == 0) { shmax[threadIdx.x >> 5] = vmax; shmin[threadIdx.x >> 5] = vmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { vmax = fmax(vmax, shmax[i]); vmin = fmin(vmin, shmin[i]); }
        partial[2 * blockIdx.x] = vmax;
        partial[2 * blockIdx.x + 1] = vmin;
    }
}

// temp = (fn <= p) * im  [+ residual]                                   (F:137-140)
__global__ void __launch_bounds__(256)
drain_threshold_kernel(const float *__restrict__ dt, const uint8_t *__restrict__ im, const double *__restrict__ pc_user,
                       const uint8_t *__restrict__ residual, int64_t n, const __grid_constant__ DrainFn q, double p,
                       uint8_t *__restrict__ temp)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const bool pore = im[v] != 0;
        bool t = pore && (drain_fn_at(q, dt[v], pore, pc_user, v) <= p);
        if (residual && residual[v]) t = true;
        temp[v] = t ? 1 : 0;
    }
}

// All pressure steps at once (ascending pressures: the sets (fn <= p_k) * im [+ residual] are nested): the class of a
// voxel is the first step k at which it belongs to the set (254: never, 255: not a node).  One flood over the
// classes (psb200_flood_classes: union-find with join times) then replaces the flood of every step.
struct DrainPs { double p[254]; };
__global__ void __launch_bounds__(256)
drain_classify_kernel(const float *__restrict__ dt, const uint8_t *__restrict__ im, const double *__restrict__ pc_user,
                      const uint8_t *__restrict__ residual, int64_t n, const __grid_constant__ DrainFn q,
                      const __grid_constant__ DrainPs ps, int np, uint8_t *__restrict__ cls)
{
    __shared__ double sp[254];
    for (int i = threadIdx.x; i < np; i += blockDim.x) sp[i] = ps.p[i];
    __syncthreads();
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const bool pore = im[v] != 0;
        uint32_t c = 255u;
        if (residual && residual[v]) c = 0u;                       // temp = ... + residual: a node at every step
        else if (pore) {
            const double f = drain_fn_at(q, dt[v], pore, pc_user, v);
            // first k with f <= p_k (p ascending): binary search; NaN compares false everywhere -> never
            int lo = 0, hi = np;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (f <= sp[mid]) hi = mid; else lo = mid + 1;
            }
            c = lo < np ? (uint32_t)lo : 254u;
        }
        cls[v] = (uint8_t)c;
    }
}

// step k of the loop from the flood of all steps: newly = (rcls == k) [* mask]; rad = int(dt) there, else 0
__global__ void __launch_bounds__(256)
drain_newly_rcls_kernel(const uint8_t *__restrict__ rcls, int k, const uint8_t *__restrict__ mask,
                        const float *__restrict__ dt, uint16_t *__restrict__ rad, int64_t n, unsigned long long *__restrict__ count,
                        int *__restrict__ maxr)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    int m = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        uint32_t r = 0;
        if ((int)rcls[v] == k && (!mask || mask[v] != 0)) {
            r = (uint32_t)min((int)dt[v], 65535);
            ++c;
            m = max(m, (int)r);
        }
        rad[v] = (uint16_t)r;
    }
    c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)c);
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if (lane_id() == 0 && c) { atomicAdd(count, c); atomicMax(maxr, m); }
}

// new_seeds = reached [* mask]; newly = new_seeds & ~seeds; seeds |= new_seeds; rad = int(dt) at newly voxels
// (F:142-152).  stats[0] += number of newly invaded voxels, stats[1] = max radius among them.
__global__ void __launch_bounds__(256)
drain_newly_kernel(const uint8_t *__restrict__ reached, const uint8_t *__restrict__ mask, uint8_t *__restrict__ seeds,
                   const float *__restrict__ dt, uint16_t *__restrict__ rad, int64_t n, unsigned long long *__restrict__ count,
                   int *__restrict__ maxr)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    int m = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step) {
        const bool ns = reached[v] != 0 && (!mask || mask[v] != 0);
        uint32_t r = 0;
        if (ns && !seeds[v]) {
            r = (uint32_t)min((int)dt[v], 65535);                      // dt[coords].astype(int)
            seeds[v] = 1;
            ++c;
            m = max(m, (int)r);
        }
        rad[v] = (uint16_t)r;
    }
    c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)c);
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if (lane_id() == 0 && c) { atomicAdd(count, c); atomicMax(maxr, m); }
}

// x pass of the power diagram: g(x) = min(C, min_x' C - rad(x')^2 + (x - x')^2) along the contiguous axis.
// One block per (line, segment of PX_SEG outputs); the segment plus r_max halo is staged in shared memory as
// f values; segments without a seed are filled with C without scanning.
#define PX_SEG 1024
__global__ void __launch_bounds__(256)
power_x_kernel(const uint16_t *__restrict__ rad, uint32_t *__restrict__ g, int64_t nlines, int nx, uint32_t C, int rmax)
{
    extern __shared__ uint32_t pxs[];                     // [PX_SEG + 2 rmax]
    const int nseg = (nx + PX_SEG - 1) / PX_SEG;
    for (int64_t job = blockIdx.x; job < nlines * nseg; job += gridDim.x) {
        const int64_t line = job / nseg;
        const int x0 = (int)(job % nseg) * PX_SEG;
        const int len = min(PX_SEG, nx - x0);
        const uint16_t *row = rad + line * nx;
        int any = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < len + 2 * rmax; i += blockDim.x) {
            const int x = x0 - rmax + i;
            uint32_t f = C;
            if (x >= 0 && x < nx) {
                const uint32_t r = row[x];
                if (r) { f = C - r * r; any = 1; }                  // r <= rmax: non-negative
            }
            pxs[i] = f;
        }
        any = __syncthreads_or(any);
        uint32_t *out = g + line * nx + x0;
        if (!any) {
            for (int i = threadIdx.x; i < len; i += blockDim.x) out[i] = C;
            continue;
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const uint32_t *c = pxs + rmax + i;
            uint32_t best = c[0];
            for (int d = 1; d <= rmax && (uint32_t)(d * d) < best; ++d)
                best = min(best, min(c[-d], c[d]) + (uint32_t)(d * d));
            out[i] = best;
        }
    }
}

// inv[v] = val where the power distance is below C and nothing was written yet   (im[x, y, z] == 0 test of
// _insert_disks_at_points, tools/_sphere_insertions.py:380-384)
__global__ void __launch_bounds__(256)
drain_paint_kernel(const uint32_t *__restrict__ e, uint8_t *__restrict__ inv, int64_t n, uint32_t C, uint32_t val)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step)
        if (e[v] < C && inv[v] == 0) inv[v] = (uint8_t)val;
}

// dst[v] = value where mask[v] != 0  (and, with `also_zero`, where dst[v] is a zero code and mask2[v] != 0)
__global__ void __launch_bounds__(256)
set_where_u8_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ mask, uint8_t value, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step)
        if (mask[v]) dst[v] = value;
}

// inv[(inv == 0) * im] = inf  (F:157): codes whose VALUE is 0 (zero_lut[code] != 0) become `value` inside the pore
__global__ void __launch_bounds__(256)
set_zero_codes_kernel(uint8_t *__restrict__ codes, const uint8_t *__restrict__ im, const uint8_t *__restrict__ zero_lut,
                      uint8_t value, int64_t n)
{
    __shared__ uint8_t z[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) z[i] = zero_lut[i];
    __syncthreads();
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += step)
        if (im[v] && z[codes[v]]) codes[v] = value;
}
