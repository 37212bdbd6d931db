/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked or called by the product path
 * (porespy_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * What it restates
 * ----------------
 * The exact Euclidean distance transform that PoreSpy obtains from the third-party
 * PyPI package `edt` (seung-lab/euclidean-distance-transform-3d; UNPINNED in
 * /root/reference/pyproject.toml:29, source not vendored under /root/reference).  Its
 * published algorithm: pass 1 along x is a two-direction linear scan, passes 2 and 3
 * along y and z take the lower envelope of parabolas (Felzenszwalb & Huttenlocher 2012 /
 * Meijster et al. 2000), threads work on independent lines, the image border is NOT
 * background (black_border=False), result = float32(sqrt(d2)) with d2 an exact integer.
 * Reference call sites on the hot path: src/porespy/filters/_funcs.py:1126 (edt(im > 0)),
 * :1191 (edt(~imtemp) < r), src/porespy/tools/_funcs.py:1153 (ps_round).
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against scipy's
 * exact EDT -- the implementation behind the shimmed reference run that reproduces the
 * reference's golden values (test/unit/test_filters.py:36-42) -- and through the
 * committed fixtures in tests/golden/.
 *
 * All arithmetic is integer (int64 intermediates); no floating point is involved until
 * the caller takes sqrt.  A line with no background voxel carries the sentinel
 * ORACLE_INF through the passes; a volume with no background at all returns ORACLE_INF
 * everywhere (the Python side maps it to +inf).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC edt_oracle.c -o liboracle.so   (oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_INF 0xFFFFFFFFu

/* pass 1: squared distance to the nearest zero along a contiguous line */
static void scan_line_x(const uint8_t *in, uint32_t *out, int64_t n)
{
    int64_t last = -1;                       /* index of the last zero seen */
    for (int64_t i = 0; i < n; i++) {
        if (in[i] == 0) { last = i; out[i] = 0; }
        else if (last < 0) out[i] = ORACLE_INF;
        else { int64_t d = i - last; out[i] = (uint32_t)(d * d); }
    }
    last = -1;
    for (int64_t i = n - 1; i >= 0; i--) {
        if (in[i] == 0) { last = i; }
        else if (last >= 0) {
            int64_t d = last - i; uint64_t dd = (uint64_t)(d * d);
            if (dd < out[i]) out[i] = (uint32_t)dd;
        }
    }
}

/* passes 2/3: out[u] = min_i (u-i)^2 + g[i] over finite g[i]; strided line.
 * Meijster's integer formulation: s[] = parabola apexes on the lower envelope,
 * t[] = first integer abscissa where s[q] takes over from s[q-1]. */
static void envelope_line(uint32_t *line, int64_t n, int64_t stride,
                          int64_t *g, int64_t *s, int64_t *t)
{
    int64_t q = -1;
    for (int64_t u = 0; u < n; u++) {
        uint32_t v = line[u * stride];
        g[u] = (v == ORACLE_INF) ? -1 : (int64_t)v;
    }
    for (int64_t u = 0; u < n; u++) {
        if (g[u] < 0) continue;              /* infinite parabola: never on the envelope */
        while (q >= 0) {
            int64_t a = t[q] - s[q], b = t[q] - u;
            if (a * a + g[s[q]] > b * b + g[u]) q--; else break;
        }
        if (q < 0) { q = 0; s[0] = u; t[0] = 0; }
        else {
            /* Sep(i,u) = floor((u^2 - i^2 + g(u) - g(i)) / (2(u-i))), i < u */
            int64_t i = s[q];
            int64_t num = u * u - i * i + g[u] - g[i], den = 2 * (u - i);
            int64_t sep = num >= 0 ? num / den : -((-num + den - 1) / den);
            int64_t w = sep + 1;
            if (w < n) { q++; s[q] = u; t[q] = w < 0 ? 0 : w; }
        }
    }
    if (q < 0) return;                       /* whole line infinite: leave as is */
    for (int64_t u = n - 1; u >= 0; u--) {
        int64_t d = u - s[q];
        line[u * stride] = (uint32_t)(d * d + g[s[q]]);
        if (u == t[q] && q > 0) q--;
    }
}

/* Exact squared EDT of the non-zero voxels of a C-contiguous [nz,ny,nx] uint8 volume
 * (nz=1 for 2-D).  out is uint32 [nz,ny,nx].  Returns 0. */
int oracle_edt_sq(const uint8_t *in, uint32_t *out, int64_t nz, int64_t ny, int64_t nx,
                  int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int64_t nlines = nz * ny;
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < nlines; l++)
        scan_line_x(in + l * nx, out + l * nx, nx);

    int64_t nmax = ny > nz ? ny : nz;
#pragma omp parallel
    {
        int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)nmax);
        int64_t *g = buf, *s = buf + nmax, *t = buf + 2 * nmax;
        if (ny > 1) {
#pragma omp for schedule(static) collapse(2)
            for (int64_t z = 0; z < nz; z++)
                for (int64_t x = 0; x < nx; x++)
                    envelope_line(out + z * ny * nx + x, ny, nx, g, s, t);
        }
        if (nz > 1) {
#pragma omp for schedule(static) collapse(2)
            for (int64_t y = 0; y < ny; y++)
                for (int64_t x = 0; x < nx; x++)
                    envelope_line(out + y * nx + x, nz, ny * nx, g, s, t);
        }
        free(buf);
    }
    return 0;
}

/* dt[i] = float32(sqrt(d2[i])) exactly as numpy's np.sqrt(float32) (IEEE, correctly
 * rounded); ORACLE_INF -> +inf.  Kept in C so the CPU baseline does not pay a numpy
 * temporary. */
#include <math.h>
int oracle_sqrt_f32(const uint32_t *d2, float *out, int64_t n, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        out[i] = d2[i] == ORACLE_INF ? INFINITY : sqrtf((float)d2[i]);
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
