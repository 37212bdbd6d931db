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
        const int64_t gu = g[u];
        if (gu < 0) continue;                /* infinite parabola: never on the envelope */
        while (q >= 0) {
            int64_t a = t[q] - s[q], b = t[q] - u;
            if (a * a + g[s[q]] > b * b + gu) q--; else break;
        }
        if (q < 0) { q = 0; s[0] = u; t[0] = 0; }
        else {
            /* Sep(i,u) = floor((u^2 - i^2 + g(u) - g(i)) / (2(u-i))), i < u */
            int64_t i = s[q];
            int64_t num = u * u - i * i + gu - g[i], den = 2 * (u - i);
            /* num / den < n  <=>  the new parabola takes over inside the line; test before dividing */
            if (num < den * (n - 1)) {
                int64_t sep = num >= 0 ? num / den : -((-num + den - 1) / den);
                int64_t w = sep + 1;
                if (w < n) { q++; s[q] = u; t[q] = w < 0 ? 0 : w; }
            }
        }
    }
    if (q < 0) return;                       /* whole line infinite: leave as is */
    for (int64_t u = n - 1; u >= 0; u--) {
        int64_t d = u - s[q];
        line[u * stride] = (uint32_t)(d * d + g[s[q]]);
        if (u == t[q] && q > 0) q--;
    }
}

/* Column passes walk COLB adjacent columns at a time through a transposed scratch block, so that
 * every cache line fetched from the strided volume is used in full (the one-column walk touches a
 * new line per element). */
#define COLB 16

static void envelope_block(uint32_t *base, int64_t n, int64_t stride, int64_t ncols,
                           uint32_t *blk, int64_t *g, int64_t *s, int64_t *t)
{
    for (int64_t u = 0; u < n; u++)
        for (int64_t c = 0; c < ncols; c++) blk[c * n + u] = base[u * stride + c];
    for (int64_t c = 0; c < ncols; c++) envelope_line(blk + c * n, n, 1, g, s, t);
    for (int64_t u = 0; u < n; u++)
        for (int64_t c = 0; c < ncols; c++) base[u * stride + c] = blk[c * n + u];
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
    int64_t nxb = (nx + COLB - 1) / COLB;
#pragma omp parallel
    {
        int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)nmax);
        uint32_t *blk = (uint32_t *)malloc(sizeof(uint32_t) * COLB * (size_t)nmax);
        int64_t *g = buf, *s = buf + nmax, *t = buf + 2 * nmax;
        if (ny > 1) {
#pragma omp for schedule(static) collapse(2)
            for (int64_t z = 0; z < nz; z++)
                for (int64_t xb = 0; xb < nxb; xb++) {
                    int64_t x = xb * COLB, nc = nx - x < COLB ? nx - x : COLB;
                    envelope_block(out + z * ny * nx + x, ny, nx, nc, blk, g, s, t);
                }
        }
        if (nz > 1) {
#pragma omp for schedule(static) collapse(2)
            for (int64_t y = 0; y < ny; y++)
                for (int64_t xb = 0; xb < nxb; xb++) {
                    int64_t x = xb * COLB, nc = nx - x < COLB ? nx - x : COLB;
                    envelope_block(out + y * nx + x, nz, ny * nx, nc, blk, g, s, t);
                }
        }
        free(buf);
        free(blk);
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

/* ------------------------------------------------------------------------------------------
 * The radius loop of `porosimetry` (mode='dt') in one C call, for volumes where the numpy
 * restatement in oracle/cpu.py would take many minutes (1024^3).  It follows
 * /root/reference/src/porespy/filters/_funcs.py line by line:
 *   F:1126  dt = edt(im > 0)
 *   F:1180  imtemp = dt >= r                      (float32 array vs the radius object: the compare
 *                                                  runs in float64 when r is float64 / int64 and in
 *                                                  float32 when r is float32 -- both are the double
 *                                                  compare below, since float32 -> double is exact)
 *   F:1181-1183  imtemp = trim_disconnected_blobs(imtemp, inlets, cross)      [access_limited]
 *   F:1184  if np.any(imtemp):
 *   F:1191      imtemp = edt(~imtemp) < r
 *   F:1192      imresults[(imresults == 0) * imtemp] = r
 * trim_disconnected_blobs (F:1265-1269: label(inlets + seeds), keep the labels that hold an inlet
 * voxel, `* im`) is evaluated as a flood from the inlet voxels through `inlets | seeds` with the
 * cross neighbourhood.  The seed sets of descending radii are nested and the inlets are fixed, so
 * the reached set only grows: the flood continues from the previous radius' reached set (new seed
 * voxels that touch a reached voxel, or are inlets themselves, start it).  tests/test_oracle.py
 * pins this call against the plain numpy restatement (oracle/cpu.py porosimetry) on small volumes.
 *
 * radii: the array the reference iterates over (descending), as doubles (float32 radii widened
 * exactly).  inlets: uint8 mask or NULL (not access-limited).  out: float64 [nz,ny,nx], zeroed here.
 */
static inline void flood_push(uint32_t *queue, int64_t *tail, uint8_t *reached, int64_t v)
{
    reached[v] = 1;
    queue[(*tail)++] = (uint32_t)v;
}

int oracle_porosimetry_dt(const uint8_t *im, const double *radii, int nr, const uint8_t *inlets,
                          double *out, int64_t nz, int64_t ny, int64_t nx, int nthreads)
{
    const int64_t n = nz * ny * nx, plane = ny * nx;
    if (inlets && n >= 0xFFFFFFFFLL) return -2;       /* the flood queue holds uint32 voxel ids */
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    uint8_t *fg = (uint8_t *)malloc((size_t)n), *seeds = (uint8_t *)malloc((size_t)n);
    uint8_t *notseeds = (uint8_t *)malloc((size_t)n);
    uint32_t *d2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    float *dt = (float *)malloc(sizeof(float) * (size_t)n);
    uint8_t *reached = NULL, *node = NULL;
    uint32_t *queue = NULL;
    if (!fg || !seeds || !notseeds || !d2 || !dt) return -1;
    if (inlets) {
        reached = (uint8_t *)calloc((size_t)n, 1);
        node = (uint8_t *)calloc((size_t)n, 1);
        queue = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        if (!reached || !node || !queue) return -1;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) { fg[i] = im[i] != 0; out[i] = 0.0; }
    oracle_edt_sq(fg, d2, nz, ny, nx, 0);
    oracle_sqrt_f32(d2, dt, n, 0);
    int first = 1;
    for (int k = 0; k < nr; k++) {
        const double r = radii[k];
        int64_t any = 0;
#pragma omp parallel for schedule(static) reduction(+ : any)
        for (int64_t i = 0; i < n; i++) { seeds[i] = (double)dt[i] >= r; any += seeds[i]; }
        if (inlets) {
            /* nodes of the graph: inlets | seeds (F:1265); new nodes start the continued flood */
            int64_t head = 0, tail = 0;
            for (int64_t i = 0; i < n; i++) {
                const uint8_t nd = (uint8_t)(seeds[i] | (inlets[i] != 0));
                if (nd && !node[i]) {
                    node[i] = 1;
                    int touch = inlets[i] != 0;
                    if (!touch && !first) {
                        const int64_t x = i % nx, y = (i / nx) % ny, z = i / plane;
                        touch = (x > 0 && reached[i - 1]) || (x + 1 < nx && reached[i + 1]) ||
                                (y > 0 && reached[i - nx]) || (y + 1 < ny && reached[i + nx]) ||
                                (z > 0 && reached[i - plane]) || (z + 1 < nz && reached[i + plane]);
                    }
                    if (touch && !reached[i]) flood_push(queue, &tail, reached, i);
                }
            }
            first = 0;
            while (head < tail) {
                const int64_t v = queue[head++];
                const int64_t x = v % nx, y = (v / nx) % ny, z = v / plane;
                if (x > 0 && node[v - 1] && !reached[v - 1]) flood_push(queue, &tail, reached, v - 1);
                if (x + 1 < nx && node[v + 1] && !reached[v + 1]) flood_push(queue, &tail, reached, v + 1);
                if (y > 0 && node[v - nx] && !reached[v - nx]) flood_push(queue, &tail, reached, v - nx);
                if (y + 1 < ny && node[v + nx] && !reached[v + nx]) flood_push(queue, &tail, reached, v + nx);
                if (z > 0 && node[v - plane] && !reached[v - plane]) flood_push(queue, &tail, reached, v - plane);
                if (z + 1 < nz && node[v + plane] && !reached[v + plane]) flood_push(queue, &tail, reached, v + plane);
            }
            any = 0;
#pragma omp parallel for schedule(static) reduction(+ : any)
            for (int64_t i = 0; i < n; i++) { seeds[i] = (uint8_t)(seeds[i] & reached[i]); any += seeds[i]; }
        }
        if (!any) continue;                                             /* F:1184 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) notseeds[i] = !seeds[i];
        oracle_edt_sq(notseeds, d2, nz, ny, nx, 0);                     /* F:1191 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            const float d = d2[i] == ORACLE_INF ? INFINITY : sqrtf((float)d2[i]);
            if (out[i] == 0.0 && (double)d < r) out[i] = r;            /* F:1192 */
        }
    }
    free(fg); free(seeds); free(notseeds); free(d2); free(dt);
    free(reached); free(node); free(queue);
    return 0;
}
