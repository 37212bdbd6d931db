// psb200.cu -- C-ABI entry points of libpsb200.so (see include/psb200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/psb200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "edt_kernels.cuh"
#include "flood_kernels.cuh"
#include "lt_kernels.cuh"
#include "minplus_kernels.cuh"
#include "xdist_kernels.cuh"
#include "bitball_kernels.cuh"
#include "blobs_kernels.cuh"
#include "sizemap_kernels.cuh"
#include "drainage_kernels.cuh"

// ------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(PSB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                     \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                        \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        CUDA_TRY(cudaGetLastError());                                                        \
    } while (0)

// ---------------------------------------------------------------------------- profiling
// Optional (ctx option "profile"): a cudaEvent pair around every kernel launch, on the
// launching stream, summed per kernel family by psb200_profile_read().
enum KernelId {
    K_EDT_X = 0, K_EDT_Y, K_EDT_Z, K_SQRT, K_MAX, K_CLASSIFY, K_LT_XY, K_LT_X, K_LT_Y, K_LT_Z, K_LT_POINT, K_EXPAND,
    K_MARK_WRITTEN, K_UF_INIT, K_UF_ACTIVATE, K_UF_MARK, K_FLOOD_MISC, K_GEN_X, K_GEN_Y, K_GEN_Z,
    K_FH_X, K_FH_Y, K_FH_Z, K_LT_PACK, K_LT_BITBALL, K_LT_WMASK, K_EDT_FIX, K_UF_FACE, K_BLOBS, K_SIZEMAP, K_DRAIN, K_COUNT
};
static const char *const kKernelNames[K_COUNT] = {
    "edt_x", "edt_y", "edt_z", "sqrt_f32", "max_u32", "lt_classify", "lt_xy", "lt_x", "lt_y", "lt_z", "lt_point",
    "lt_expand", "lt_mark_written", "uf_init", "uf_activate", "uf_mark", "flood_misc",
    "generic_x", "generic_y", "generic_z", "edt_fh_x", "edt_fh_y", "edt_fh_z", "lt_pack", "lt_bitball", "lt_wmask", "edt_fix_inf",
    "uf_face", "blobs", "sizemap", "drainage"};

struct ProfScope {
    psb200_ctx *c;
    cudaStream_t st;
    int kid;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(psb200_ctx *c_, cudaStream_t st_, int kid_) : c(c_), st(st_), kid(kid_)
    {
        if (!c->profile) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, st);
    }
    ~ProfScope()
    {
        if (!c->profile || !a || !b) return;
        cudaEventRecord(b, st);
        c->prof.push_back(ProfRec{kid, a, b});
    }
};

extern "C" int psb200_version(void) { return PSB200_VERSION; }
extern "C" const char *psb200_last_error(void) { return g_err; }

extern "C" int psb200_create(int device, psb200_ctx **out)
{
    if (!out) return fail(PSB200_ERR_INVALID, "psb200_create: ctx pointer is NULL");
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(PSB200_ERR_INVALID, "psb200_create: device %d not in [0,%d)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    psb200_ctx *c = new psb200_ctx();
    c->device = device;
    c->algo = PSB200_ALGO_FAST;
    c->launches = 0;
    c->profile = 0;
    c->bit_tmax = 200;
    c->edt16 = 1;
    c->bit4 = 1;
    c->foot = 1;
    c->ycoarse = 1;
    c->xbits = 1;
    c->bitquad = 1;
    c->ydirect = 1;
    c->edt_h = 32;
    c->zwide = 1;
    c->uf_records = 1;
    c->yflags = 1;
    c->flag_slot = 0;
    CUDA_TRY(cudaMalloc(&c->flags, 64 * sizeof(int)));
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaFuncSetAttribute(lt_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(lt_y2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(lt_y2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(lt_y3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(xdist_kernel<XD_EDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(xdist_kernel<XD_LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus_kernel<MpSrcU16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus_kernel<MpSrcU16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus_kernel<MpSrcU32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus_kernel<MpSrcU32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU16, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU16, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU16, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU32, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU32, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU32, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_minplus16_kernel<MpSrcU32, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_x_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    CUDA_TRY(cudaFuncSetAttribute(edt_x_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    *out = c;
    return PSB200_OK;
}

extern "C" int psb200_destroy(psb200_ctx *ctx)
{
    if (ctx && ctx->flags) {
        cudaSetDevice(ctx->device);
        cudaFree(ctx->flags);
    }
    delete ctx;
    return PSB200_OK;
}

extern "C" int psb200_set_option(psb200_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return fail(PSB200_ERR_INVALID, "set_option: NULL argument");
    if (!strcmp(name, "algo")) {
        if (value != PSB200_ALGO_FAST && value != PSB200_ALGO_GENERIC)
            return fail(PSB200_ERR_INVALID, "set_option: algo must be 0 (fast) or 1 (generic)");
        ctx->algo = (int)value;
        return PSB200_OK;
    }
    if (!strcmp(name, "bit_tmax")) {
        if (value < 0 || value > 400) return fail(PSB200_ERR_INVALID, "set_option: bit_tmax must be in [0,400]");
        ctx->bit_tmax = (int)value;
        return PSB200_OK;
    }
    if (!strcmp(name, "bit4")) {
        ctx->bit4 = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "uf_records")) {
        ctx->uf_records = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "yflags")) {
        ctx->yflags = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "zwide")) {
        ctx->zwide = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "edt_h")) {
        if (value < 8 || value > 48 || (value & 3)) return fail(PSB200_ERR_INVALID, "set_option: edt_h must be a multiple of 4 in [8,48]");
        ctx->edt_h = (int)value;
        return PSB200_OK;
    }
    if (!strcmp(name, "ydirect")) {
        ctx->ydirect = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "bitquad")) {
        ctx->bitquad = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "xbits")) {
        ctx->xbits = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "ycoarse")) {
        ctx->ycoarse = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "foot")) {
        ctx->foot = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "edt16")) {
        ctx->edt16 = value ? 1 : 0;
        return PSB200_OK;
    }
    if (!strcmp(name, "profile")) {
        ctx->profile = value ? 1 : 0;
        return PSB200_OK;
    }
    return fail(PSB200_ERR_INVALID, "set_option: unknown option '%s'", name);
}

extern "C" int64_t psb200_launch_count(const psb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int psb200_profile_kernels(void) { return K_COUNT; }
extern "C" const char *psb200_profile_name(int kernel_id)
{
    return (kernel_id >= 0 && kernel_id < K_COUNT) ? kKernelNames[kernel_id] : "";
}

// Per-launch records in launch order (kernel id, milliseconds); returns the number of records
// available (<= max written).  Does not clear.
extern "C" int psb200_profile_records(psb200_ctx *ctx, int *kernel_ids, float *ms, int max)
{
    if (!ctx) return 0;
    int i = 0;
    for (ProfRec &r : ctx->prof) {
        if (i < max && kernel_ids && ms) {
            if (cudaEventSynchronize(r.b) != cudaSuccess) break;
            kernel_ids[i] = r.kid;
            if (cudaEventElapsedTime(&ms[i], r.a, r.b) != cudaSuccess) ms[i] = -1.f;
        }
        ++i;
    }
    return i;
}

// Synchronises on the recorded events.  ms_total / launches: arrays of psb200_profile_kernels()
// entries (summed per kernel family since the last reset).  Clears the records.
extern "C" int psb200_profile_read(psb200_ctx *ctx, double *ms_total, int64_t *launches)
{
    if (!ctx || !ms_total || !launches) return fail(PSB200_ERR_INVALID, "profile_read: NULL argument");
    for (int i = 0; i < K_COUNT; ++i) { ms_total[i] = 0.0; launches[i] = 0; }
    for (ProfRec &r : ctx->prof) {
        CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_total[r.kid] += ms;
        launches[r.kid] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    return PSB200_OK;
}

// --------------------------------------------------------------------------- utilities
static int check_dims(const char *who, int64_t nz, int64_t ny, int64_t nx)
{
    if (nz < 1 || ny < 1 || nx < 1)
        return fail(PSB200_ERR_INVALID, "%s: empty volume (%lld,%lld,%lld)", who, (long long)nz,
                    (long long)ny, (long long)nx);
    if (nz > PSB200_MAX_DIM || ny > PSB200_MAX_DIM || nx > PSB200_MAX_DIM)
        return fail(PSB200_ERR_UNSUPPORTED, "%s: dimension exceeds %d", who, PSB200_MAX_DIM);
    return PSB200_OK;
}

struct Carver {
    char *base;
    size_t off;
    template <typename T>
    T *take(size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

static int grid_for(int64_t n, int block, int sm_count, int per_sm)
{
    int64_t g = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

#define COL_BLOCK 128
#define COL_BLOCKS_PER_SM 8

static int64_t col_threads(const psb200_ctx *ctx) { return (int64_t)ctx->sm_count * COL_BLOCKS_PER_SM * COL_BLOCK; }

static size_t stack_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny)
{
    int64_t nmax = ny > nz ? ny : nz;
    if (ny <= 1 && nz <= 1) return 0;
    return (size_t)col_threads(ctx) * (size_t)nmax * sizeof(uint2);
}

// ---------------------------------------------------------------------------------- EDT
// Fast path: xdist (u16) -> bounded min-plus y pass -> bounded min-plus z pass.
// Workspace: [u16 x-distances: n][u32 intermediate: n (only when both ny > 1 and nz > 1)].
// ALGO_GENERIC: the lower-envelope (Felzenszwalb / Meijster) kernels, in place on d2.
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t psb200_edt_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx)
{
    if (!ctx) return 0;
    const size_t n = (size_t)nz * ny * nx;
    // fast path: uint16 x-distances, uint32 y-pass result, the scalar, and the stacks of its envelope fallback
    return align256(n * 2) + align256(n * 4) + 256 + align256(stack_bytes(ctx, nz, ny)) + 1024;
}

template <int SITE_MODE>
static int launch_x(psb200_ctx *ctx, const uint8_t *in, uint32_t *d2, int64_t nlines, int nx, int k,
                    cudaStream_t st)
{
    const int nwords = (nx + 31) / 32;
    int warps = 8;
    while (warps > 1 && (size_t)warps * 3 * nwords * 4 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 3 * nwords * 4;
    const int grid = grid_for(nlines, warps, ctx->sm_count, 16);
    {
        ProfScope ps__(ctx, st, SITE_MODE == 0 ? K_FH_X : K_GEN_X);
        edt_x_kernel<SITE_MODE><<<grid, warps * 32, smem, st>>>(in, d2, nlines, nx, k);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// axis 1 (y) or 0 (z), in place on d2 unless a Store functor redirects the output
template <typename Store>
static int launch_col(psb200_ctx *ctx, int axis, const uint32_t *src, Store store, int64_t nz,
                      int64_t ny, int64_t nx, void *ws, size_t ws_bytes, cudaStream_t st,
                      int prof_kid = -1, const int *gate = nullptr)
{
    if (prof_kid < 0) prof_kid = axis == 1 ? K_FH_Y : K_FH_Z;
    const int64_t plane = ny * nx;
    int64_t ncols, inner, outer, stride;
    int n;
    if (axis == 1) { ncols = nz * nx; inner = nx; outer = plane; stride = nx; n = (int)ny; }
    else { ncols = plane; inner = plane; outer = 0; stride = plane; n = (int)nz; }
    const int64_t need = (int64_t)col_threads(ctx) * n * (int64_t)sizeof(uint2);
    if (!ws || (int64_t)ws_bytes < need)
        return fail(PSB200_ERR_WORKSPACE, "column pass needs %lld workspace bytes, got %lld",
                    (long long)need, (long long)ws_bytes);
    const int grid = grid_for(ncols, COL_BLOCK, ctx->sm_count, COL_BLOCKS_PER_SM);
    {
        ProfScope ps__(ctx, st, prof_kid);
        edt_col_kernel<Store><<<grid, COL_BLOCK, 0, st>>>(src, store, ncols, inner, outer, stride, n,
                                                          reinterpret_cast<uint2 *>(ws), gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// x-distance pass (xdist_kernel): one warp per line
template <int MODE>
static int launch_xdist(psb200_ctx *ctx, const uint8_t *in, void *out, int64_t nlines, int nx, int k, int cap,
                        const int *gate, cudaStream_t st)
{
    const int nch = (nx + 15) / 16;
    int warps = 8;
    while (warps > 1 && (size_t)warps * 2 * nch * 4 > 64 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 2 * nch * 4;
    const int grid = grid_for(nlines, warps, ctx->sm_count, 32);
    {
        ProfScope ps__(ctx, st, MODE == XD_EDT ? K_EDT_X : K_LT_X);
        xdist_kernel<MODE><<<grid, warps * 32, smem, st>>>(in, out, nlines, nx, k, cap, gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static bool ntiles_check(int64_t gx, int64_t nouter) { return gx > 0 && nouter > (int64_t)0x7FFFFFFFFFFFLL / gx; }

// bounded min-plus pass along y (axis 1) or z (axis 0); out_kind 0: u32 squared, 1: f32 sqrt
template <typename Src>
static int launch_minplus(psb200_ctx *ctx, int axis, const typename Src::T *src, void *dst, int out_kind,
                          uint32_t *gmax, int64_t nz, int64_t ny, int64_t nx, cudaStream_t st, int split = 0,
                          int mrow0 = 0, int mrow1 = 0x7FFFFFFF, int **ovf_out = nullptr)
{
    // ovf_out != nullptr: the caller provides the fallback for overflowed 16-bit passes itself (edt_fast: the
    // lower-envelope kernels, whose cost does not depend on the distances); *ovf_out = the device flag
    const int64_t plane = ny * nx;
    int n;
    int64_t rstride, nxc, ostride, nouter;
    if (axis == 1) { n = (int)ny; rstride = nx; nxc = nx; ostride = plane; nouter = nz; }
    else { n = (int)nz; rstride = plane; nxc = plane; ostride = 0; nouter = 1; }
    int L, H;
    if (n <= 128) { L = n; H = 0; }
    else { L = 64; H = 32; }
    const size_t smem = (size_t)(((L + 3) & ~3) + 2 * H) * 512;
    const int vec = (nxc % 4 == 0) && (rstride % 4 == 0) && (ostride % 4 == 0) &&
                    ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0);
    const int64_t gx = ((nxc + MP_TX - 1) / MP_TX) * ((n + L - 1) / L);
    if (gx > 0x7FFFFFFFLL || nouter > 65535)
        return fail(PSB200_ERR_UNSUPPORTED, "edt pass: volume too large for one launch");
    if (ntiles_check(gx, nouter)) return fail(PSB200_ERR_UNSUPPORTED, "edt pass: too many tiles");
    const int64_t ntiles = gx * nouter;
    // one CTA holds (L + 2H) * 512 bytes of shared memory: 3 resident per SM; the blocks walk the tiles
    const unsigned grid = (unsigned)(ntiles < (int64_t)ctx->sm_count * 3 ? ntiles : (int64_t)ctx->sm_count * 3);
    const int *gate = nullptr;
    if (ctx->edt16) {
        // 16-bit form first (exact wherever the result is < 32767); the uint32 kernel behind it only
        // runs when a voxel overflowed
        int *ovf = ctx->flags + (ctx->flag_slot++ & 63u);
        int L16, H16;
        if (n <= 224) { L16 = n; H16 = 0; }
        else { L16 = 128; H16 = ctx->edt_h; }
        const size_t smem16 = (size_t)(((L16 + 3) & ~3) + 2 * H16) * MP_TS * 8 + (size_t)(H16 + 2) * 16;
        const int64_t gx16 = ((nxc + MP_TX - 1) / MP_TX) * ((n + L16 - 1) / L16);
        if (gx16 > 0x7FFFFFFFLL) return fail(PSB200_ERR_UNSUPPORTED, "edt pass: volume too large for one launch");
        dim3 grid16((unsigned)gx16, (unsigned)nouter);
        CUDA_TRY(cudaMemsetAsync(ovf, 0, sizeof(int), st));
        {
            ProfScope ps__(ctx, st, axis == 1 ? K_EDT_Y : K_EDT_Z);
            if (out_kind == 0 && !ctx->foot)
                edt_minplus16_kernel<Src, 0, 0><<<grid16, MP16_WARPS * 32, smem16, st>>>(src, dst, n, rstride, nxc, ostride, L16, H16, vec, gmax, split, ovf, mrow0, mrow1);
            else if (out_kind == 0)
                edt_minplus16_kernel<Src, 0, 1><<<grid16, MP16_WARPS * 32, smem16, st>>>(src, dst, n, rstride, nxc, ostride, L16, H16, vec, gmax, split, ovf, mrow0, mrow1);
            else if (!ctx->foot)
                edt_minplus16_kernel<Src, 1, 0><<<grid16, MP16_WARPS * 32, smem16, st>>>(src, dst, n, rstride, nxc, ostride, L16, H16, vec, gmax, split, ovf, mrow0, mrow1);
            else
                edt_minplus16_kernel<Src, 1, 1><<<grid16, MP16_WARPS * 32, smem16, st>>>(src, dst, n, rstride, nxc, ostride, L16, H16, vec, gmax, split, ovf, mrow0, mrow1);
        }
        LAUNCH_CHECK(ctx);
        gate = ovf;
        if (ovf_out) { *ovf_out = ovf; return PSB200_OK; }
    }
    if (ovf_out) *ovf_out = nullptr;
    {
        ProfScope ps__(ctx, st, axis == 1 ? K_EDT_Y : K_EDT_Z);
        if (out_kind == 0)
            edt_minplus_kernel<Src, 0><<<grid, MP_WARPS * 32, smem, st>>>(src, dst, n, rstride, nxc, ostride, L, H, vec, gmax, split, gate, gx, (int)nouter, mrow0, mrow1);
        else
            edt_minplus_kernel<Src, 1><<<grid, MP_WARPS * 32, smem, st>>>(src, dst, n, rstride, nxc, ostride, L, H, vec, gmax, split, gate, gx, (int)nouter, mrow0, mrow1);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// last pass epilogue: map infinite values to PSB_INF (device-gated on the running max)
static int launch_fix_inf(psb200_ctx *ctx, void *out, int out_kind, uint32_t *gmax, int64_t n, cudaStream_t st)
{
    {
        ProfScope ps__(ctx, st, K_EDT_FIX);
        edt_fix_inf_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(
            out_kind == 0 ? reinterpret_cast<uint32_t *>(out) : nullptr, n, gmax);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// Gated fallback of one overflowed 16-bit pass: the lower-envelope kernel on uint32 data in place.  A pass overflows
// when some squared distance along the axes done so far reaches 32767 (181 voxels) -- never on porous media, but a
// volume with a single background voxel has distances of a thousand voxels, and a bounded min-plus scan (O(distance)
// per voxel, from global memory beyond the staged halo) then takes seconds where the envelope (O(1) per voxel)
// takes tens of milliseconds.  Every kernel below returns at once when *ovf == 0.
static int edt_envelope_fallback(psb200_ctx *ctx, int axis, uint32_t *data, void *out, int out_kind, uint32_t *gmax,
                                 int64_t nz, int64_t ny, int64_t nx, int zmax0, int zmax1, bool last, void *stk,
                                 size_t stk_bytes, const int *ovf, cudaStream_t st)
{
    const int64_t n = nz * ny * nx, plane = ny * nx;
    int rc = launch_col(ctx, axis, data, EdtStoreU32{data}, nz, ny, nx, stk, stk_bytes, st, axis == 1 ? K_EDT_Y : K_EDT_Z, ovf);
    if (rc || !last) return rc;
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    if (gmax) {
        const int64_t z0 = nz > 1 ? zmax0 : 0, z1 = nz > 1 ? (zmax1 < nz ? zmax1 : nz) : 1;
        {
            ProfScope ps__(ctx, st, K_MAX);
            max_u32_kernel<<<g, 256, 0, st>>>(data + z0 * plane, (z1 - z0) * plane, gmax, ovf);
        }
        LAUNCH_CHECK(ctx);
    }
    if (out_kind == 1) {
        ProfScope ps__(ctx, st, K_SQRT);
        sqrt_f32_kernel<<<g, 256, 0, st>>>(data, reinterpret_cast<float *>(out), n, ovf);
        LAUNCH_CHECK(ctx);
    }
    return PSB200_OK;
}

static int edt_fast(psb200_ctx *ctx, const uint8_t *in, void *out, int out_kind, uint32_t *gmax,
                    int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes, cudaStream_t st,
                    int zmax0 = 0, int zmax1 = 0x7FFFFFFF)
{
    const size_t n = (size_t)nz * ny * nx;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    const size_t off_mid = align256(n * 2), off_max = off_mid + ((nz > 1) ? align256(n * 4) : 0), off_stk = off_max + 256;
    // the envelope fallback serves passes of more than one row; a one-row pass keeps the (trivial) uint32 min-plus one
    const bool env_y = ctx->edt16 && ny > 1, env_z = ctx->edt16 && nz > 1;
    const size_t stk_bytes = (env_y || env_z) ? stack_bytes(ctx, nz, ny) : 0;
    const size_t need = off_stk + stk_bytes + 256;
    if (!base || ws_bytes < need + (size_t)(base - (char *)ws))
        return fail(PSB200_ERR_WORKSPACE, "edt needs %zu workspace bytes, got %zu", need + 256, ws_bytes);
    uint16_t *dx = reinterpret_cast<uint16_t *>(base);
    uint32_t *mid = reinterpret_cast<uint32_t *>(base + off_mid);
    void *stk = base + off_stk;
    if (!gmax) gmax = reinterpret_cast<uint32_t *>(base + off_max);
    CUDA_TRY(cudaMemsetAsync(gmax, 0, sizeof(uint32_t), st));
    const int g = grid_for((int64_t)n, 256, ctx->sm_count, 16);
    int rc = launch_xdist<XD_EDT>(ctx, in, dx, nz * ny, (int)nx, 0, 0, nullptr, st);
    if (rc) return rc;
    int *ovf = nullptr;
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);          // float32 and uint32 have the same size
    if (nz == 1) {
        rc = launch_minplus<MpSrcU16>(ctx, 1, dx, out, out_kind, gmax, nz, ny, nx, st, 0, 0, 0x7FFFFFFF, env_y ? &ovf : nullptr);
        if (rc) return rc;
        if (ovf) {
            sq16_to_u32_kernel<<<g, 256, 0, st>>>(dx, out32, (int64_t)n, ovf);
            LAUNCH_CHECK(ctx);
            rc = edt_envelope_fallback(ctx, 1, out32, out, out_kind, gmax, nz, ny, nx, 0, 1, true, stk, stk_bytes, ovf, st);
        }
    } else {
        rc = launch_minplus<MpSrcU16>(ctx, 1, dx, mid, 0, nullptr, nz, ny, nx, st, 0, 0, 0x7FFFFFFF, env_y ? &ovf : nullptr);
        if (rc) return rc;
        if (ovf) {
            sq16_to_u32_kernel<<<g, 256, 0, st>>>(dx, mid, (int64_t)n, ovf);
            LAUNCH_CHECK(ctx);
            rc = edt_envelope_fallback(ctx, 1, mid, nullptr, 0, nullptr, nz, ny, nx, 0, 0, false, stk, stk_bytes, ovf, st);
            if (rc) return rc;
        }
        ovf = nullptr;
        rc = launch_minplus<MpSrcU32>(ctx, 0, mid, out, out_kind, gmax, nz, ny, nx, st, 0, zmax0, zmax1, env_z ? &ovf : nullptr);
        if (rc) return rc;
        if (ovf) {
            copy_u32_kernel<<<g, 256, 0, st>>>(mid, out32, (int64_t)n, ovf);
            LAUNCH_CHECK(ctx);
            rc = edt_envelope_fallback(ctx, 0, out32, out, out_kind, gmax, nz, ny, nx, zmax0, zmax1, true, stk, stk_bytes, ovf, st);
        }
    }
    if (rc) return rc;
    return launch_fix_inf(ctx, out, out_kind, gmax, (int64_t)n, st);
}

extern "C" int psb200_edt_pass(psb200_ctx *ctx, int axis, const uint8_t *in, uint32_t *d2,
                               int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                               psb200_stream stream)
{
    if (!ctx || !d2) return fail(PSB200_ERR_INVALID, "edt_pass: NULL argument");
    int rc = check_dims("edt_pass", nz, ny, nx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    void *wsa = ws ? (void *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    size_t wsb = ws ? ws_bytes - ((uintptr_t)wsa - (uintptr_t)ws) : 0;
    if (axis == 2) {
        if (!in) return fail(PSB200_ERR_INVALID, "edt_pass: x pass needs the input image");
        return launch_x<0>(ctx, in, d2, nz * ny, (int)nx, 0, st);
    }
    if (axis == 1) {
        if (ny == 1) return PSB200_OK;
        return launch_col(ctx, 1, d2, EdtStoreU32{d2}, nz, ny, nx, wsa, wsb, st);
    }
    if (axis == 0) {
        if (nz == 1) return PSB200_OK;
        return launch_col(ctx, 0, d2, EdtStoreU32{d2}, nz, ny, nx, wsa, wsb, st);
    }
    return fail(PSB200_ERR_INVALID, "edt_pass: axis must be 0, 1 or 2");
}

extern "C" int psb200_edt_u8(psb200_ctx *ctx, const uint8_t *in, void *out, int out_kind,
                             uint32_t *max_out, int64_t nz, int64_t ny, int64_t nx, void *ws,
                             size_t ws_bytes, psb200_stream stream)
{
    return psb200_edt_u8_zmax(ctx, in, out, out_kind, max_out, nz, ny, nx, 0, nz, ws, ws_bytes, stream);
}

extern "C" int psb200_edt_u8_zmax(psb200_ctx *ctx, const uint8_t *in, void *out, int out_kind,
                                  uint32_t *max_out, int64_t nz, int64_t ny, int64_t nx, int64_t zmax0, int64_t zmax1,
                                  void *ws, size_t ws_bytes, psb200_stream stream)
{
    if (!ctx || !in || !out) return fail(PSB200_ERR_INVALID, "edt_u8: NULL argument");
    if (zmax0 < 0 || zmax1 > nz || zmax0 > zmax1) return fail(PSB200_ERR_INVALID, "edt_u8_zmax: plane range outside [0, nz]");
    const bool ranged = !(zmax0 == 0 && zmax1 == nz);
    if (ranged && (nz == 1 || ctx->algo == PSB200_ALGO_GENERIC))
        return fail(PSB200_ERR_UNSUPPORTED, "edt_u8_zmax: a plane range needs a 3-D volume and the fast algorithm");
    if (out_kind != 0 && out_kind != 1) return fail(PSB200_ERR_INVALID, "edt_u8: out_kind must be 0 (u32 d2) or 1 (f32)");
    int rc = check_dims("edt_u8", nz, ny, nx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->algo == PSB200_ALGO_GENERIC) {
        // lower-envelope kernels; u32 in place, then the pointwise epilogues
        uint32_t *d2 = reinterpret_cast<uint32_t *>(out);      // float32 and uint32 have the same size
        for (int axis = 2; axis >= 0; --axis) {
            rc = psb200_edt_pass(ctx, axis, in, d2, nz, ny, nx, ws, ws_bytes, stream);
            if (rc) return rc;
        }
        const int64_t n = nz * ny * nx;
        if (max_out) {
            rc = psb200_max_u32(ctx, d2, n, max_out, stream);
            if (rc) return rc;
        }
        if (out_kind == 1) return psb200_sqrt_f32(ctx, d2, reinterpret_cast<float *>(out), n, stream);
        return PSB200_OK;
    }
    return edt_fast(ctx, in, out, out_kind, max_out, nz, ny, nx, ws, ws_bytes, st, (int)zmax0, (int)zmax1);
}

extern "C" int psb200_edt_sq_u8(psb200_ctx *ctx, const uint8_t *in, uint32_t *d2, int64_t nz,
                                int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                                psb200_stream stream)
{
    return psb200_edt_u8(ctx, in, d2, 0, nullptr, nz, ny, nx, ws, ws_bytes, stream);
}

// ---- z-slab sharded EDT: x and y passes on the local slab, z pass on the pencil
extern "C" int psb200_edt_xy_u8(psb200_ctx *ctx, const uint8_t *in, uint32_t *h_out, int64_t nz,
                                int64_t ny, int64_t nx, int64_t ysplit, void *ws, size_t ws_bytes,
                                psb200_stream stream)
{
    if (!ctx || !in || !h_out) return fail(PSB200_ERR_INVALID, "edt_xy_u8: NULL argument");
    int rc = check_dims("edt_xy_u8", nz, ny, nx);
    if (rc) return rc;
    if (ysplit < 0 || ysplit > ny) return fail(PSB200_ERR_INVALID, "edt_xy_u8: ysplit outside [0, ny]");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)nz * ny * nx;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    if (!base || ws_bytes < align256(n * 2) + 512)
        return fail(PSB200_ERR_WORKSPACE, "edt_xy_u8 needs %zu workspace bytes, got %zu", align256(n * 2) + 512, ws_bytes);
    uint16_t *dx = reinterpret_cast<uint16_t *>(base);
    rc = launch_xdist<XD_EDT>(ctx, in, dx, nz * ny, (int)nx, 0, 0, nullptr, st);
    if (rc) return rc;
    return launch_minplus<MpSrcU16>(ctx, 1, dx, h_out, 0, nullptr, nz, ny, nx, st, (int)ysplit);
}

extern "C" int psb200_edt_z_u32(psb200_ctx *ctx, const uint32_t *h, void *out, int out_kind,
                                uint32_t *max_out, int64_t nz, int64_t ny, int64_t nx,
                                psb200_stream stream)
{
    if (!ctx || !h || !out || h == out) return fail(PSB200_ERR_INVALID, "edt_z_u32: bad argument");
    if (out_kind != 0 && out_kind != 1) return fail(PSB200_ERR_INVALID, "edt_z_u32: out_kind must be 0 or 1");
    int rc = check_dims("edt_z_u32", nz, ny, nx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!max_out) return fail(PSB200_ERR_INVALID, "edt_z_u32: max_out is required (device uint32)");
    CUDA_TRY(cudaMemsetAsync(max_out, 0, sizeof(uint32_t), st));
    rc = launch_minplus<MpSrcU32>(ctx, 0, h, out, out_kind, max_out, nz, ny, nx, st);
    if (rc) return rc;
    return launch_fix_inf(ctx, out, out_kind, max_out, nz * ny * nx, st);
}

extern "C" int psb200_sqrt_f32(psb200_ctx *ctx, const uint32_t *d2, float *out, int64_t n,
                               psb200_stream stream)
{
    if (!ctx || !d2 || !out || n < 0) return fail(PSB200_ERR_INVALID, "sqrt_f32: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        ProfScope ps__(ctx, (cudaStream_t)stream, K_SQRT);
        sqrt_f32_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, (cudaStream_t)stream>>>(d2, out, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_max_u32(psb200_ctx *ctx, const uint32_t *d2, int64_t n, uint32_t *dev_out,
                              psb200_stream stream)
{
    if (!ctx || !d2 || !dev_out || n < 0) return fail(PSB200_ERR_INVALID, "max_u32: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(dev_out, 0, sizeof(uint32_t), st));
    if (n == 0) return PSB200_OK;
    {
        ProfScope ps__(ctx, st, K_MAX);
        max_u32_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(d2, n, dev_out);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ------------------------------------------------------------------ local thickness loop
static uint32_t isqrt_u32(uint32_t v)
{
    uint32_t r = (uint32_t)__builtin_sqrt((double)v);
    while ((uint64_t)r * r > v) --r;
    while ((uint64_t)(r + 1) * (r + 1) <= v) ++r;
    return r;
}

static int check_thresholds(const char *who, const uint32_t *T, int nT)
{
    if (nT < 0 || nT > PSB200_MAX_THRESHOLDS)
        return fail(PSB200_ERR_INVALID, "%s: nT=%d outside [0,%d]", who, nT, PSB200_MAX_THRESHOLDS);
    if (nT && !T) return fail(PSB200_ERR_INVALID, "%s: thresholds pointer is NULL", who);
    for (int k = 0; k < nT; ++k) {
        if (T[k] == 0) return fail(PSB200_ERR_INVALID, "%s: threshold %d is 0", who, k);
        if (k && T[k] >= T[k - 1])
            return fail(PSB200_ERR_INVALID, "%s: thresholds must be strictly descending", who);
    }
    return PSB200_OK;
}

// access-limited flooding: the voxels that become active at one radius are listed segment by
// segment (UF_SEG voxels of the volume per segment bound the list), then linked by one thread
// per (voxel, neighbour) job
#define UF_SEG (1LL << 27)
static size_t uf_list_entries(int64_t n) { return (size_t)(n < UF_SEG ? n : UF_SEG); }

struct LtWorkspace {
    uint8_t *cls, *rcls, *reach, *gx;
    uint8_t *xflag;                 // byte path: activity byte per 32-voxel word from the bit-based x pass
    uint32_t *seedbits, *written;   // bit path: one bit per voxel (seed bits of up to PACKN_MAX consecutive radii)
    size_t seed_words;              // words of one seed-bit volume inside seedbits
    uint32_t *parent;
    uint32_t *uf_list;    // link records sorted into (radius index, direction) slices; job list of the fallback path
    size_t uf_list_cap;
    uint32_t *uf_raw;     // unsorted records, one region of uf_cap entries per chunk; uf_ccount: records per chunk
    uint32_t *uf_ccount;
    int64_t uf_chunks;
    uint32_t uf_cap;
    uint32_t *uf_bins;    // histogram / slice starts / cursors of the link records
    int *gate;
    uint32_t *gen_d2;     // generic algo: full u32 distance map of ~seeds
    char *gen_stk;
    size_t gen_stk_bytes;
    size_t total;
};

static LtWorkspace carve_lt(const psb200_ctx *ctx, char *base, int64_t nz, int64_t ny, int64_t nx,
                            int inlet_mode)
{
    const size_t n = (size_t)nz * ny * nx;
    Carver c{base, 0};
    LtWorkspace w{};
    w.gate = c.take<int>(1024);     // [0] gate, [1] first reached radius, [64..] class histogram / slice starts / cursors
    w.cls = c.take<uint8_t>(n + 16);
    w.reach = c.take<uint8_t>(n + 16);
    w.gx = c.take<uint8_t>(n + 16);
    w.xflag = c.take<uint8_t>(n / 32 + 64);
    w.seed_words = (n / 32 + 4 + 63) & ~(size_t)63;                 // keeps every volume 256-byte aligned
    w.seedbits = c.take<uint32_t>(w.seed_words * PACKN_MAX);
    w.written = c.take<uint32_t>(n / 32 + 4);
    if (inlet_mode != PSB200_INLETS_NONE) {
        w.rcls = c.take<uint8_t>(n + 16);
        w.parent = c.take<uint32_t>(n + 1);
        // link records of the row-rooted forest (flood_kernels.cuh): per chunk of UF_CHUNK segments at most one x
        // record per voxel and one per two voxels (+ one per segment) for each of y and z -- raw[] has that much
        // room per chunk, list[] (the records sorted into slices; also the job list of the fallback path) in total
        const int64_t segs = nz * ny * ((nx + UF_SEGX - 1) / UF_SEGX);
        w.uf_chunks = (segs + UF_CHUNK - 1) / UF_CHUNK;
        w.uf_cap = (uint32_t)(2 * UF_CHUNK * (nx < UF_SEGX ? nx : UF_SEGX) + 2 * UF_CHUNK + 64);
        w.uf_list_cap = std::max<size_t>(n + 64, (size_t)w.uf_chunks * w.uf_cap);
        w.uf_list = c.take<uint32_t>(w.uf_list_cap);
        w.uf_raw = c.take<uint32_t>((size_t)w.uf_chunks * w.uf_cap);
        w.uf_ccount = c.take<uint32_t>((size_t)w.uf_chunks + 64);
        w.uf_bins = c.take<uint32_t>(3 * (UF_NTIMES * UF_MAXFAM + 64));
    }
    if (ctx->algo == PSB200_ALGO_GENERIC) {
        w.gen_d2 = c.take<uint32_t>(n);
        w.gen_stk_bytes = stack_bytes(ctx, nz, ny);
        w.gen_stk = c.take<char>(w.gen_stk_bytes);
    }
    w.total = c.off + 256;
    return w;
}

extern "C" size_t psb200_local_thickness_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny,
                                                         int64_t nx, int inlet_mode)
{
    if (!ctx) return 0;
    return carve_lt(ctx, nullptr, nz, ny, nx, inlet_mode).total + 256;
}

extern "C" size_t psb200_flood_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx)
{
    if (!ctx) return 0;
    psb200_ctx tmp = *ctx;
    tmp.algo = PSB200_ALGO_FAST;
    return carve_lt(&tmp, nullptr, nz, ny, nx, PSB200_INLETS_MASK).total + 256;
}

static int classify_impl(psb200_ctx *ctx, const uint32_t *d2, const uint32_t *T_host, int nT,
                         uint8_t *cls, int64_t n, cudaStream_t st)
{
    TArg targ;
    memset(&targ, 0, sizeof(targ));
    for (int k = 0; k < nT; ++k) targ.t[k] = T_host[k];
    {
        ProfScope ps__(ctx, st, K_CLASSIFY);
        lt_classify_kernel<<<grid_for((n + 3) / 4, 256, ctx->sm_count, 16), 256, 0, st>>>(d2, cls, n, targ, nT);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_lt_classify(psb200_ctx *ctx, const uint32_t *d2, const uint32_t *T_host, int nT,
                                  uint8_t *cls, int64_t n, psb200_stream stream)
{
    if (!ctx || !d2 || !cls || n < 0) return fail(PSB200_ERR_INVALID, "lt_classify: bad argument");
    int rc = check_thresholds("lt_classify", T_host, nT);
    if (rc) return rc;
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return classify_impl(ctx, d2, T_host, nT, cls, n, (cudaStream_t)stream);
}

static int lt_xy_impl(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t T, uint8_t *reach,
                      int64_t nz, int64_t ny, int64_t nx, const int *gate, cudaStream_t st)
{
    const int W = (int)isqrt_u32(T - 1);
    if (W > LT_MAX_W)
        return fail(PSB200_ERR_UNSUPPORTED, "lt_xy: threshold %u exceeds the uint8 pipeline (r > 254)", T);
    int Ly = ny < 64 ? (int)ny : 64;
    const int HW = (W + 31) / 32, NW = 4 + 2 * HW;
    size_t smem = (size_t)(Ly + 2 * W) * LT_XT + (size_t)LT_WARPS * NW * 4;
    if ((int)smem > ctx->max_smem_optin)
        return fail(PSB200_ERR_UNSUPPORTED, "lt_xy: tile needs %zu bytes of shared memory", smem);
    dim3 grid((unsigned)((nx + LT_XT - 1) / LT_XT), (unsigned)((ny + Ly - 1) / Ly), (unsigned)nz);
    {
        ProfScope ps__(ctx, st, K_LT_XY);
        lt_xy_kernel<<<grid, LT_WARPS * 32, smem, st>>>(cls, reach, (int)ny, (int)nx, k, T, W, Ly, gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

#define LT_LZ 32
#define LTY3_MAX_W 26

static int lt_z_impl(psb200_ctx *ctx, const uint8_t *reach, const uint8_t *m_lo, int nlo,
                     const uint8_t *m_hi, int nhi, uint8_t *idx, int k, uint32_t T, int64_t nz,
                     int64_t ny, int64_t nx, const int *gate, cudaStream_t st)
{
    const int W = (int)isqrt_u32(T - 1);
    const int64_t plane = ny * nx;
    if (nlo > W) { m_lo += (int64_t)(nlo - W) * plane; nlo = W; }
    if (nhi > W) nhi = W;
    const bool vec4 = (plane % 4 == 0) && (((uintptr_t)reach | (uintptr_t)idx | (uintptr_t)m_lo | (uintptr_t)m_hi) % 4 == 0);
    const unsigned gy = (unsigned)((nz + LT_LZ - 1) / LT_LZ);
    if (vec4) {
        dim3 grid((unsigned)((plane / 4 + 255) / 256), gy);
        {
            ProfScope ps__(ctx, st, K_LT_Z);
            lt_z_kernel<LT_LZ, 4><<<grid, 256, 0, st>>>(reach, m_lo, nlo, m_hi, nhi, idx, (int)nz, plane, W,
                                                       (uint32_t)(k + 1), gate);
        }
    } else {
        dim3 grid((unsigned)((plane + 255) / 256), gy);
        {
            ProfScope ps__(ctx, st, K_LT_Z);
            lt_z_kernel<LT_LZ, 1><<<grid, 256, 0, st>>>(reach, m_lo, nlo, m_hi, nhi, idx, (int)nz, plane, W,
                                                       (uint32_t)(k + 1), gate);
        }
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// streaming three-kernel form (nx % 16 == 0, T <= 32767): x-distance, y tile scan, in-place z sweeps
static bool streaming_ok(int64_t ny, int64_t nx, uint32_t T, const void *a, const void *b, const void *c)
{
    (void)ny;
    return (nx % 16 == 0) && T <= 32767u && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15u) == 0);
}

static int lt_xy_stream_impl(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t T, uint8_t *gx,
                             uint8_t *reach, int64_t nz, int64_t ny, int64_t nx, const int *gate,
                             cudaStream_t st, const uint32_t *seedbits = nullptr, uint8_t *xflag = nullptr)
{
    const int W = (int)isqrt_u32(T - 1);
    int rc = PSB200_OK;
    const bool from_bits = seedbits && nx % 32 == 0 && ctx->xbits;
    if (!from_bits || !ctx->yflags) xflag = nullptr;
    if (from_bits) {
        // the seed set of this radius is already packed: x pass from the bits (xdist_bits_kernel)
        const int nw = (int)(nx / 32);
        int warps = 8;
        while (warps > 1 && (size_t)warps * 2 * nw * 4 > 48 * 1024) warps >>= 1;
        {
            ProfScope ps__(ctx, st, K_LT_X);
            xdist_bits_kernel<<<grid_for(nz * ny, warps, ctx->sm_count, 32), warps * 32, (size_t)warps * 2 * nw * 4, st>>>(
                seedbits, gx, nz * ny, nw, W + 1, gate, xflag);
        }
        LAUNCH_CHECK(ctx);
    } else
        rc = launch_xdist<XD_LT>(ctx, cls, gx, nz * ny, (int)nx, k, W + 1, gate, st);
    if (rc) return rc;
    {   // y pass
        int Ly = ny < 128 ? (int)ny : 128;
        // hierarchical scan for the denser radii (measured r2c at 1024^3: it wins from W = 25 down -- 2.49 -> 2.25 ms
        // at T = 344 -- and loses above, where seeds are sparse and its larger halo and group minima only cost)
        const int direct = ctx->ydirect;
        const bool coarse = ctx->ycoarse && W <= LTY3_MAX_W && (int)lt_y3_smem_bytes(Ly, W, T, direct) <= ctx->max_smem_optin;
        const size_t smem = coarse ? lt_y3_smem_bytes(Ly, W, T, direct) : lt_y2_smem_bytes(Ly, W, T, direct);
        if ((int)smem > ctx->max_smem_optin)
            return fail(PSB200_ERR_UNSUPPORTED, "lt_y: tile needs %zu bytes of shared memory", smem);
        dim3 grid((unsigned)((nx + MP_TX - 1) / MP_TX), (unsigned)((ny + Ly - 1) / Ly), (unsigned)nz);
        {
            ProfScope ps__(ctx, st, K_LT_Y);
            // (the 32 x 16 warp footprint that helps the EDT passes is 1 % slower here: r2b)
            if (coarse) lt_y3_kernel<<<grid, 256, smem, st>>>(gx, reach, (int)ny, (int)nx, T, W, Ly, gate, direct);
            else lt_y2_kernel<0><<<grid, 256, smem, st>>>(gx, reach, (int)ny, (int)nx, T, W, Ly, gate, direct, xflag);
        }
        LAUNCH_CHECK(ctx);
    }
    return PSB200_OK;
}

static int lt_z_stream_impl(psb200_ctx *ctx, uint8_t *reach, const uint8_t *m_lo, int nlo,
                            const uint8_t *m_hi, int nhi, uint8_t *idx, int k, uint32_t T, int64_t nz,
                            int64_t ny, int64_t nx, const int *gate, cudaStream_t st)
{
    const int W = (int)isqrt_u32(T - 1);
    const int64_t plane = ny * nx;
    if (nlo > W) { m_lo += (int64_t)(nlo - W) * plane; nlo = W; }
    if (nhi > W) nhi = W;
    const bool wide = ctx->zwide && plane % 8 == 0 &&
                      ((((uintptr_t)reach | (uintptr_t)idx | (uintptr_t)m_lo | (uintptr_t)m_hi) & 7u) == 0);
    {
        ProfScope ps__(ctx, st, K_LT_Z);
        if (wide)
            lt_zsweep8_kernel<<<(unsigned)((plane / 8 + 127) / 128), 128, 0, st>>>(reach, m_lo, nlo, m_hi, nhi, idx, (int)nz, plane,
                                                                                 (uint32_t)(k + 1), gate);
        else
            lt_zsweep_kernel<<<(unsigned)((plane / 4 + 255) / 256), 256, 0, st>>>(reach, m_lo, nlo, m_hi, nhi, idx, (int)nz, plane,
                                                                                (uint32_t)(k + 1), gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static int lt_pack_impl(psb200_ctx *ctx, const uint8_t *cmap, int k, uint32_t *bits, int64_t nwords,
                        const int *gate, cudaStream_t st);

extern "C" int psb200_lt_xy(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t T, uint8_t *reach,
                            int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                            psb200_stream stream)
{
    if (!ctx || !cls || !reach || T == 0 || k < 0 || k >= PSB200_MAX_THRESHOLDS)
        return fail(PSB200_ERR_INVALID, "lt_xy: bad argument");
    int rc = check_dims("lt_xy", nz, ny, nx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint8_t *gx = ws ? (uint8_t *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    const size_t n = (size_t)(nz * ny * nx);
    const size_t need = n + 256, used = gx ? (size_t)(gx - (uint8_t *)ws) : 0;
    if (gx && ws_bytes >= need && streaming_ok(ny, nx, T, cls, reach, gx)) {
        // with room for one bit per voxel behind the x-distance bytes: pack the seeds of this radius and take the x
        // pass from the bits (lt_pack 0.2 ms + xdist_bits 0.4 ms instead of xdist 0.75 ms at 1024^3)
        const size_t off_bits = align256(n + 16);
        if (ctx->xbits && nx % 32 == 0 && ws_bytes >= used + off_bits + n / 8 + 256 && ((uintptr_t)cls & 15u) == 0) {
            uint32_t *bits = reinterpret_cast<uint32_t *>(gx + off_bits);
            const size_t off_flag = off_bits + align256(n / 8 + 16);
            uint8_t *xflag = ws_bytes >= used + off_flag + n / 32 + 256 ? gx + off_flag : nullptr;
            int rc2 = lt_pack_impl(ctx, cls, k, bits, (int64_t)(n / 32), nullptr, (cudaStream_t)stream);
            if (rc2) return rc2;
            return lt_xy_stream_impl(ctx, cls, k, T, gx, reach, nz, ny, nx, nullptr, (cudaStream_t)stream, bits, xflag);
        }
        return lt_xy_stream_impl(ctx, cls, k, T, gx, reach, nz, ny, nx, nullptr, (cudaStream_t)stream);
    }
    return lt_xy_impl(ctx, cls, k, T, reach, nz, ny, nx, nullptr, (cudaStream_t)stream);
}

extern "C" int psb200_lt_z(psb200_ctx *ctx, uint8_t *reach, const uint8_t *m_lo, int nlo,
                           const uint8_t *m_hi, int nhi, uint8_t *idx, int k, uint32_t T, int64_t nz,
                           int64_t ny, int64_t nx, psb200_stream stream)
{
    if (!ctx || !reach || !idx || T == 0 || k < 0 || k >= PSB200_MAX_THRESHOLDS || nlo < 0 || nhi < 0 ||
        (nlo && !m_lo) || (nhi && !m_hi))
        return fail(PSB200_ERR_INVALID, "lt_z: bad argument");
    int rc = check_dims("lt_z", nz, ny, nx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (streaming_ok(ny, nx, 1, reach, idx, nullptr) && ((((uintptr_t)m_lo | (uintptr_t)m_hi) & 3u) == 0))
        return lt_z_stream_impl(ctx, reach, m_lo, nlo, m_hi, nhi, idx, k, T, nz, ny, nx, nullptr,
                                (cudaStream_t)stream);
    return lt_z_impl(ctx, reach, m_lo, nlo, m_hi, nhi, idx, k, T, nz, ny, nx, nullptr, (cudaStream_t)stream);
}

extern "C" int psb200_lt_halo_cone(psb200_ctx *ctx, const uint8_t *reach, int64_t nz, int64_t ny, int64_t nx, int depth,
                                   int side, uint8_t *out, psb200_stream stream)
{
    if (!ctx || !reach || !out || depth < 0 || (side != 0 && side != 1)) return fail(PSB200_ERR_INVALID, "lt_halo_cone: bad argument");
    int rc = check_dims("lt_halo_cone", nz, ny, nx);
    if (rc) return rc;
    if (depth > nz) depth = (int)nz;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t plane = ny * nx;
    {
        ProfScope ps__(ctx, st, K_LT_Z);
        lt_halo_cone_kernel<<<grid_for((plane + 3) / 4, 256, ctx->sm_count, 8), 256, 0, st>>>(reach, (int)nz, plane, depth, side, out);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static int uf_activate_impl(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const InletSpec &inl, int klo,
                            int khi, int conn, int64_t nz, int64_t ny, int64_t nx, uint32_t *list,
                            cudaStream_t st, uint8_t *jtime = nullptr)
{
    const int64_t n = nz * ny * nx;
    uint32_t *count = list;                 // entry 0..15: the counter, the list starts at entry 16
    uint32_t *items = list + 16;
    for (int64_t v0 = 0; v0 < n; v0 += UF_SEG) {
        const int64_t v1 = v0 + UF_SEG < n ? v0 + UF_SEG : n;
        CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint32_t), st));
        {
            ProfScope ps__(ctx, st, K_UF_ACTIVATE);
            uf_collect_kernel<<<grid_for((v1 - v0 + 63) / 64, 256, ctx->sm_count, 8), 256, 0, st>>>(
                cls, klo, khi, v0, v1, items, count);
        }
        LAUNCH_CHECK(ctx);
        {
            ProfScope ps__(ctx, st, K_UF_ACTIVATE);
            uf_union_list_kernel<<<ctx->sm_count * 16, 256, 0, st>>>(parent, cls, inl, klo, khi, conn, (int)nz,
                                                                     (int)ny, (int)nx, items, count, jtime, nullptr);
        }
        LAUNCH_CHECK(ctx);
    }
    return PSB200_OK;
}

// Forest initialisation with the x chains pre-linked, and the link records of all radius indices bucketed by
// (index, direction) -- see flood_kernels.cuh.  `acls` receives the activation map (class, inlets folded to 0).
// Returns the device array of slice starts through *start_out.  conn 26: the record count has no useful a-priori
// bound, so the host reads the total and *fits tells whether the records were written.
static int uf_forest_impl(psb200_ctx *ctx, LtWorkspace &w, const InletSpec &inl, const uint8_t *cls, uint8_t *acls,
                          uint8_t *jtime, int conn, int64_t nz, int64_t ny, int64_t nx, cudaStream_t st,
                          const uint32_t **start_out, bool *fits)
{
    const int nfam = conn == 6 ? 3 : UF_MAXFAM, nsub = conn == 6 ? 2 * nfam : nfam;
    const int nbins = UF_NTIMES * nsub;
    const int nseg = (int)((nx + UF_SEGX - 1) / UF_SEGX);
    const int64_t segs = nz * ny * nseg;
    const int grid_a = (int)std::min<int64_t>((segs + 7) / 8, (int64_t)ctx->sm_count * 8);
    const int grid = (int)std::min<int64_t>((segs + UF_CHUNK - 1) / UF_CHUNK, (int64_t)ctx->sm_count * 8);
    uint32_t *hist = w.uf_bins, *start = hist + nbins + 64, *cursor = start + nbins + 64;
    {
        ProfScope ps__(ctx, st, K_UF_INIT);
        uf_prelink_kernel<<<grid_a, 256, 0, st>>>(cls, inl, (int)nz, (int)ny, (int)nx, w.parent, jtime, acls);
    }
    LAUNCH_CHECK(ctx);
    int *ovf = w.gate + 2;
    CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(ovf, 0, sizeof(int), st));
    {
        ProfScope ps__(ctx, st, K_UF_ACTIVATE);
        if (nfam > 3)
            uf_emit_kernel<5><<<grid, 256, (size_t)nbins * 4, st>>>(acls, inl, (int)nz, (int)ny, (int)nx, nfam, nsub, hist,
                                                                  w.uf_raw, w.uf_ccount, w.uf_cap, ovf);
        else
            uf_emit_kernel<3><<<grid, 256, (size_t)nbins * 4, st>>>(acls, inl, (int)nz, (int)ny, (int)nx, nfam, nsub, hist,
                                                                  w.uf_raw, w.uf_ccount, w.uf_cap, ovf);
    }
    LAUNCH_CHECK(ctx);
    {
        ProfScope ps__(ctx, st, K_UF_ACTIVATE);
        uf_scan_bins_kernel<<<1, 1024, 0, st>>>(hist, start, cursor, nbins);
    }
    LAUNCH_CHECK(ctx);
    *fits = true;
    if (nfam > 3) {
        // 26-connectivity has no useful a-priori bound on the records of a chunk
        int over = 0;
        CUDA_TRY(cudaMemcpyAsync(&over, ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (over) { *fits = false; return PSB200_OK; }
    }
    {
        ProfScope ps__(ctx, st, K_UF_ACTIVATE);
        uf_sort_kernel<<<grid, 256, (size_t)nbins * 8, st>>>(w.uf_raw, w.uf_ccount, w.uf_cap, w.uf_chunks, nseg, (int)nx, nbins,
                                                           start, cursor, w.uf_list);
    }
    LAUNCH_CHECK(ctx);
    *start_out = start;
    return PSB200_OK;
}

static int uf_union_records(psb200_ctx *ctx, LtWorkspace &w, const uint32_t *start, int k, int conn, int64_t ny,
                            int64_t nx, uint8_t *jtime, cudaStream_t st)
{
    const int nfam = conn == 6 ? 3 : UF_MAXFAM, nsub = conn == 6 ? 2 * nfam : nfam;
    static const int8_t fam[UF_MAXFAM][3] = {{0, 0, 1},  {0, 1, 0},  {1, 0, 0},  {0, 1, -1}, {0, 1, 1},
                                             {1, 0, -1}, {1, 0, 1},  {1, 1, -1}, {1, 1, 0},  {1, 1, 1},
                                             {1, -1, -1}, {1, -1, 0}, {1, -1, 1}};
    UfStrides sd;
    for (int f = 0; f < UF_MAXFAM; ++f) sd.s[f] = ((long long)fam[f][0] * ny + fam[f][1]) * nx + fam[f][2];
    {
        ProfScope ps__(ctx, st, K_UF_ACTIVATE);
        uf_union_rec_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(w.parent, w.uf_list, start, k, nfam, nsub, sd, jtime);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static int uf_step(psb200_ctx *ctx, LtWorkspace &w, const InletSpec &inl, int klo, int khi, int conn,
                   int64_t nz, int64_t ny, int64_t nx, cudaStream_t st)
{
    const int64_t n = nz * ny * nx;
    int rc = uf_activate_impl(ctx, w.parent, w.cls, inl, klo, khi, conn, nz, ny, nx, w.uf_list, st);
    if (rc) return rc;
    {
        ProfScope ps__(ctx, st, K_UF_MARK);
        uf_mark_kernel<<<grid_for((n + 15) / 16, 256, ctx->sm_count, 16), 256, 0, st>>>(w.parent, w.cls, w.rcls, khi, n, w.gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}


// (dy,dz) offsets of the digital ball {o : |o|^2 < T}, grouped by x-allowance (bitball_kernels.cuh)
static bool build_ball_pairs(uint32_t T, int64_t ny, int64_t nw, BallPairs &bp)
{
    const int W = (int)isqrt_u32(T - 1);
    if (W > 31) return false;
    bp.W = W;
    int cnt = 0;
    for (int a = W; a >= 0; --a) {
        for (int dz = -W; dz <= W; ++dz)
            for (int dy = -W; dy <= W; ++dy) {
                const int64_t rem = (int64_t)T - 1 - (int64_t)dy * dy - (int64_t)dz * dz;
                if (rem < 0) continue;
                if ((int)isqrt_u32((uint32_t)rem) != a) continue;
                if (cnt >= BB_MAX_PAIRS) return false;
                bp.e[cnt].x = (int)((dz * ny + dy) * nw);
                bp.e[cnt].y = (int)(((uint32_t)dy & 0xFFFFu) | ((uint32_t)dz << 16));
                ++cnt;
            }
        // pad the ring to a multiple of 4 pairs with copies of its last pair (cnt stays a multiple of 4
        // across rings, so an empty ring needs nothing)
        while (cnt % 4 != 0) {
            if (cnt >= BB_MAX_PAIRS) return false;
            bp.e[cnt] = bp.e[cnt - 1];
            ++cnt;
        }
        bp.ring_end[a] = (unsigned short)cnt;
    }
    bp.ring_end[W + 1] = 0;
    return true;
}

// load lists of the two-output dilation kernel (bitball_kernels.cuh, lt_bitball4d_kernel): for every Horner stage a
// the source rows (sy, sz), relative to the upper output row, whose allowance is exactly a for both outputs
// (allow(sy, sz) == allow(sy - 1, sz) == a), for the upper one only, for the lower one only; every list is padded
// to a multiple of 4 with copies of its last entry (OR is idempotent)
static bool build_ball_duo(uint32_t T, int64_t ny, int64_t nw, BallDuo &bd)
{
    const int W = (int)isqrt_u32(T - 1);
    if (W > 31) return false;
    bd.W = W;
    int cnt = 0;
    auto allow = [&](int dy, int dz) -> int {
        const int64_t rem = (int64_t)T - 1 - (int64_t)dy * dy - (int64_t)dz * dz;
        return rem < 0 ? -1 : (int)isqrt_u32((uint32_t)rem);
    };
    for (int a = W; a >= 0; --a) {
        for (int list = 0; list < 3; ++list) {
            const int first = cnt;
            for (int sz = -W; sz <= W; ++sz)
                for (int sy = -W; sy <= W + 1; ++sy) {
                    const bool up = allow(sy, sz) == a, lo = allow(sy - 1, sz) == a;
                    const bool take = list == 0 ? (up && lo) : (list == 1 ? (up && !lo) : (!up && lo));
                    if (!take) continue;
                    if (cnt >= BB3_MAX_ENTRIES) return false;
                    bd.off[cnt++] = (int)((sz * ny + sy) * nw);
                }
            while (cnt > first && cnt % 4 != 0) {
                if (cnt >= BB3_MAX_ENTRIES) return false;
                bd.off[cnt] = bd.off[cnt - 1];
                ++cnt;
            }
            bd.end[a][list] = (unsigned short)cnt;
        }
    }
    return true;
}

static int lt_pack_impl(psb200_ctx *ctx, const uint8_t *cmap, int k, uint32_t *bits, int64_t nwords,
                        const int *gate, cudaStream_t st)
{
    {
        ProfScope ps__(ctx, st, K_LT_PACK);
        lt_pack_kernel<<<grid_for(nwords, 256, ctx->sm_count, 16), 256, 0, st>>>(cmap, bits, nwords, k, gate);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static int lt_bitball_impl(psb200_ctx *ctx, const uint32_t *seedbits, int64_t nz_src, int64_t z_off,
                           uint32_t *written, uint8_t *idx, int k, uint32_t T, int64_t nz, int64_t ny,
                           int64_t nx, const int *gate, cudaStream_t st)
{
    static thread_local BallPairs bp;
    memset(&bp, 0, sizeof(int) + sizeof(bp.ring_end));
    if (!build_ball_pairs(T, ny, nx / 32, bp)) return fail(PSB200_ERR_UNSUPPORTED, "bit path: threshold %u too large", T);
    const int nw = (int)(nx / 32);
    // rows longer than 32 words: two words per lane when the row offsets stay 8-byte aligned
    const bool two = nw > 32 && (nw % 2 == 0) && ((reinterpret_cast<uintptr_t>(seedbits) & 7u) == 0);
    const int seg = two ? (nw <= 64 ? 64 : 60) : (nw <= 32 ? 32 : 30);
    dim3 grid((unsigned)((nw + seg - 1) / seg), (unsigned)((ny + BB_TY - 1) / BB_TY), (unsigned)((nz + BB_TZ - 1) / BB_TZ));
    // rows of 32 / 64 / 128 words: four words per lane (16-byte loads)
    const bool four = ctx->bit4 && (nw == 32 || nw == 64 || nw == 128) &&
                      ((((uintptr_t)seedbits | (uintptr_t)written) & 15u) == 0);
    static thread_local BallDuo bd;
    if (four && ctx->bitquad && ny >= 2 && T >= 20 && build_ball_duo(T, ny, nx / 32, bd)) {
        // two output rows per lane: 79 % of the loads per output (small balls: list padding eats the gain)
        const int lpr = nw / 4, gz = (256 / lpr) / 8;
        dim3 gq(1, (unsigned)((ny + 15) / 16), (unsigned)((nz + gz - 1) / gz));
        {
            ProfScope ps__(ctx, st, K_LT_BITBALL);
            if (lpr == 8)
                lt_bitball4d_kernel<8><<<gq, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bd, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
            else if (lpr == 16)
                lt_bitball4d_kernel<16><<<gq, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bd, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
            else
                lt_bitball4d_kernel<32><<<gq, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bd, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
        }
        LAUNCH_CHECK(ctx);
        return PSB200_OK;
    }
    if (four) {
        const int lpr = nw / 4, tz = 32 / lpr;
        dim3 g4(1, (unsigned)((ny + 7) / 8), (unsigned)((nz + tz - 1) / tz));
        {
            ProfScope ps__(ctx, st, K_LT_BITBALL);
            if (lpr == 8)
                lt_bitball4_kernel<8><<<g4, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
            else if (lpr == 16)
                lt_bitball4_kernel<16><<<g4, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
            else
                lt_bitball4_kernel<32><<<g4, 256, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, bp, (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
        }
        LAUNCH_CHECK(ctx);
        return PSB200_OK;
    }
    {
        ProfScope ps__(ctx, st, K_LT_BITBALL);
        if (two)
            lt_bitball2_kernel<<<grid, 1024, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, nw, seg, bp,
                                                      (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
        else
            lt_bitball_kernel<<<grid, 1024, 0, st>>>(seedbits, written, idx, (int)nz, (int)ny, nw, seg, bp,
                                                     (uint32_t)(k + 1), gate, (int)nz_src, (int)z_off);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static int lt_wmask_impl(psb200_ctx *ctx, const uint8_t *idx, uint32_t *written, int64_t nwords, cudaStream_t st)
{
    {
        ProfScope ps__(ctx, st, K_LT_WMASK);
        lt_wmask_kernel<<<grid_for(nwords, 256, ctx->sm_count, 16), 256, 0, st>>>(idx, written, nwords);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

static void launch_packn(const uint8_t *cmap, uint32_t *bits, int64_t nwords, int64_t vol_words, int k0, int nk, int grid,
                         cudaStream_t st)
{
    const int kmax = k0 + nk - 1;
    if (kmax < 32) lt_packn_kernel<5><<<grid, 256, 0, st>>>(cmap, bits, nwords, vol_words, k0, nk);
    else if (kmax < 64) lt_packn_kernel<6><<<grid, 256, 0, st>>>(cmap, bits, nwords, vol_words, k0, nk);
    else if (kmax < 128) lt_packn_kernel<7><<<grid, 256, 0, st>>>(cmap, bits, nwords, vol_words, k0, nk);
    else lt_packn_kernel<8><<<grid, 256, 0, st>>>(cmap, bits, nwords, vol_words, k0, nk);
}

// Seed bits of radius k: the seed bits of radii k .. k + PACKN_MAX - 1 are packed together from one read of the
// class map (lt_packn_kernel) whenever k is not in the buffer.  `packed_lo/hi`: the radii whose bits sit in
// w.seedbits.  Both pipelines use them: the bit path dilates them, the byte path takes its x pass from them.
static int lt_seed_bits(psb200_ctx *ctx, LtWorkspace &w, const uint8_t *cmap, int k, int nT, int64_t nwords,
                        cudaStream_t st, int &packed_lo, int &packed_hi, const uint32_t **bits)
{
    if (k < packed_lo || k >= packed_hi) {
        const int nk = nT - k < PACKN_MAX ? nT - k : PACKN_MAX;
        {
            ProfScope ps__(ctx, st, K_LT_PACK);
            // not gated: the buffers must be valid at the later radii even if radius k is still before the breakthrough
            launch_packn(cmap, w.seedbits, nwords, (int64_t)w.seed_words, k, nk, grid_for(nwords, 256, ctx->sm_count, 8), st);
        }
        LAUNCH_CHECK(ctx);
        packed_lo = k;
        packed_hi = k + nk;
    }
    *bits = w.seedbits + (size_t)(k - packed_lo) * w.seed_words;
    return PSB200_OK;
}

// One radius of the bit path (thresholds descend, so every radius after the first bit radius takes it too).
static int lt_bit_step(psb200_ctx *ctx, LtWorkspace &w, const uint8_t *cmap, uint8_t *idx, int k, int nT, uint32_t T,
                       int64_t nz, int64_t ny, int64_t nx, const int *gate, cudaStream_t st, int &packed_lo,
                       int &packed_hi)
{
    const uint32_t *bits = nullptr;
    int rc = lt_seed_bits(ctx, w, cmap, k, nT, nz * ny * nx / 32, st, packed_lo, packed_hi, &bits);
    if (rc) return rc;
    return lt_bitball_impl(ctx, bits, nz, 0, w.written, idx, k, T, nz, ny, nx, gate, st);
}

// step-level entry points of the bit path (z-slab shards exchange seed-bit halo planes between them)
static int bit_shape_ok(const char *who, int64_t nx, const void *a, const void *b)
{
    if (nx % 32 != 0) return fail(PSB200_ERR_UNSUPPORTED, "%s: the bit path needs nx %% 32 == 0", who);
    if ((((uintptr_t)a | (uintptr_t)b) & 15u) != 0) return fail(PSB200_ERR_INVALID, "%s: buffers must be 16-byte aligned", who);
    return PSB200_OK;
}

extern "C" int psb200_lt_pack(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t *bits, int64_t nz,
                              int64_t ny, int64_t nx, psb200_stream stream)
{
    if (!ctx || !cls || !bits || k < 0 || k >= PSB200_MAX_THRESHOLDS) return fail(PSB200_ERR_INVALID, "lt_pack: bad argument");
    int rc = check_dims("lt_pack", nz, ny, nx);
    if (!rc) rc = bit_shape_ok("lt_pack", nx, cls, bits);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return lt_pack_impl(ctx, cls, k, bits, nz * ny * nx / 32, nullptr, (cudaStream_t)stream);
}

extern "C" int psb200_lt_packn(psb200_ctx *ctx, const uint8_t *cls, int k0, int nk, uint32_t *bits, int64_t vol_words,
                               int64_t nz, int64_t ny, int64_t nx, psb200_stream stream)
{
    if (!ctx || !cls || !bits || k0 < 0 || nk < 1 || nk > PACKN_MAX || k0 + nk > PSB200_MAX_THRESHOLDS)
        return fail(PSB200_ERR_INVALID, "lt_packn: bad argument (1 <= nk <= %d)", PACKN_MAX);
    int rc = check_dims("lt_packn", nz, ny, nx);
    if (!rc) rc = bit_shape_ok("lt_packn", nx, cls, bits);
    if (rc) return rc;
    const int64_t nwords = nz * ny * nx / 32;
    if (vol_words < nwords || (vol_words & 3)) return fail(PSB200_ERR_INVALID, "lt_packn: vol_words must be >= the words of one volume and a multiple of 4");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps__(ctx, st, K_LT_PACK);
        launch_packn(cls, bits, nwords, vol_words, k0, nk, grid_for(nwords, 256, ctx->sm_count, 8), st);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_lt_wmask(psb200_ctx *ctx, const uint8_t *idx, uint32_t *written, int64_t nz, int64_t ny,
                               int64_t nx, psb200_stream stream)
{
    if (!ctx || !idx || !written) return fail(PSB200_ERR_INVALID, "lt_wmask: NULL argument");
    int rc = check_dims("lt_wmask", nz, ny, nx);
    if (!rc) rc = bit_shape_ok("lt_wmask", nx, idx, written);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return lt_wmask_impl(ctx, idx, written, nz * ny * nx / 32, (cudaStream_t)stream);
}

extern "C" int psb200_lt_bitball(psb200_ctx *ctx, const uint32_t *seedbits, int64_t nz_src, int64_t z_off,
                                 uint32_t *written, uint8_t *idx, int k, uint32_t T, int64_t nz, int64_t ny,
                                 int64_t nx, psb200_stream stream)
{
    if (!ctx || !seedbits || !written || !idx || T == 0 || k < 0 || k >= PSB200_MAX_THRESHOLDS)
        return fail(PSB200_ERR_INVALID, "lt_bitball: bad argument");
    if (z_off < 0 || z_off + nz > nz_src) return fail(PSB200_ERR_INVALID, "lt_bitball: slab [z_off, z_off+nz) outside the seed buffer");
    int rc = check_dims("lt_bitball", nz, ny, nx);
    if (!rc) rc = bit_shape_ok("lt_bitball", nx, seedbits, idx);
    if (rc) return rc;
    if (nz_src > 65535 || nz_src * ny * nx >= (1LL << 35)) return fail(PSB200_ERR_UNSUPPORTED, "lt_bitball: volume too large");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return lt_bitball_impl(ctx, seedbits, nz_src, z_off, written, idx, k, T, nz, ny, nx, nullptr, (cudaStream_t)stream);
}

extern "C" int psb200_local_thickness_idx(psb200_ctx *ctx, const uint32_t *d2, const uint32_t *T_host,
                                          int nT, uint8_t *idx, const uint8_t *inlets, int inlet_mode,
                                          int ndim, int64_t nz, int64_t ny, int64_t nx, int flags,
                                          void *ws, size_t ws_bytes, psb200_stream stream)
{
    if (!ctx || !d2 || !idx) return fail(PSB200_ERR_INVALID, "local_thickness_idx: NULL argument");
    int rc = check_dims("local_thickness_idx", nz, ny, nx);
    if (rc) return rc;
    rc = check_thresholds("local_thickness_idx", T_host, nT);
    if (rc) return rc;
    if (inlet_mode < PSB200_INLETS_NONE || inlet_mode > PSB200_INLETS_MASK)
        return fail(PSB200_ERR_INVALID, "local_thickness_idx: bad inlet_mode %d", inlet_mode);
    if (inlet_mode == PSB200_INLETS_MASK && !inlets)
        return fail(PSB200_ERR_INVALID, "local_thickness_idx: inlet mask is NULL");
    if (ndim < 1 || ndim > 3) return fail(PSB200_ERR_INVALID, "local_thickness_idx: ndim must be 1..3");
    const int64_t n = nz * ny * nx;
    if (inlet_mode != PSB200_INLETS_NONE && n > 0xFFFFFFF0LL)
        return fail(PSB200_ERR_UNSUPPORTED, "access-limited flooding supports < 2^32 voxels per GPU");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    LtWorkspace w = carve_lt(ctx, base, nz, ny, nx, inlet_mode);
    if (!ws || w.total + 256 > ws_bytes)
        return fail(PSB200_ERR_WORKSPACE, "local_thickness_idx needs %zu workspace bytes, got %zu",
                    w.total + 256, ws_bytes);
    if (!(flags & PSB200_FLAG_IDX_PREINIT)) CUDA_TRY(cudaMemsetAsync(idx, 0, (size_t)n, st));
    if (nT == 0) return PSB200_OK;

    rc = classify_impl(ctx, d2, T_host, nT, w.cls, n, st);
    if (rc) return rc;

    const bool al = inlet_mode != PSB200_INLETS_NONE;
    InletSpec inl{inlet_mode, ndim, inlets, 0, (int)nz};
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    if (al) {
        // F:1181-1183 for every radius at once: the seed sets are nested, so the union-find links
        // the voxels of class k at step k and keeps, for every tree that joins the inlets, the step
        // at which it did (jtime, in the reach buffer: the dilation passes only start afterwards);
        // one resolve pass then yields  rcls = first radius at which the voxel is a reached seed.
        uint8_t *jtime = w.reach;
        int *kmin = w.gate + 1;
        CUDA_TRY(cudaMemsetAsync(w.gate, 0, sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(kmin, 0x7F, sizeof(int), st));
        if (ctx->uf_records) {
            // until the resolve pass the rcls buffer holds the activation map (class, inlets folded to 0)
            const uint32_t *start = nullptr;
            bool fits = true;
            rc = uf_forest_impl(ctx, w, inl, w.cls, w.rcls, jtime, 6, nz, ny, nx, st, &start, &fits);
            if (rc) return rc;
            for (int k = 0; k < nT; ++k) {
                rc = uf_union_records(ctx, w, start, k, 6, ny, nx, jtime, st);
                if (rc) return rc;
            }
            {
                ProfScope ps__(ctx, st, K_UF_MARK);
                uf_compress_kernel<<<ctx->sm_count * 6, 256, 0, st>>>(w.parent, w.rcls, (int)nz, (int)ny, (int)nx);
            }
            LAUNCH_CHECK(ctx);
        } else {
            {
                ProfScope ps__(ctx, st, K_UF_INIT);
                uf_init_kernel<<<g, 256, 0, st>>>(w.parent, inl, (int)nz, (int)ny, (int)nx, jtime, w.cls, w.rcls);
            }
            LAUNCH_CHECK(ctx);
            // until the resolve pass the rcls buffer holds the activation map (class, inlets folded to 0)
            const uint8_t *acls = w.rcls;
            const InletSpec folded{3, ndim, nullptr, 0, (int)nz};
            uint32_t *hist = reinterpret_cast<uint32_t *>(w.gate) + 64, *start = hist + 256, *cursor = start + 320;
            CUDA_TRY(cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), st));
            {
                ProfScope ps__(ctx, st, K_UF_ACTIVATE);
                uf_hist_kernel<<<grid_for(n / 16 + 1, 256, ctx->sm_count, 8), 256, 0, st>>>(acls, n, hist);
            }
            LAUNCH_CHECK(ctx);
            {
                ProfScope ps__(ctx, st, K_UF_ACTIVATE);
                uf_scan_kernel<<<1, 256, 0, st>>>(hist, start, cursor);
            }
            LAUNCH_CHECK(ctx);
            {
                ProfScope ps__(ctx, st, K_UF_ACTIVATE);
                uf_bucket_kernel<<<grid_for((n + 63) / 64, 256, ctx->sm_count, 8), 256, 0, st>>>(acls, n, start, cursor, w.uf_list);
            }
            LAUNCH_CHECK(ctx);
            for (int k = 0; k < nT; ++k) {
                {
                    ProfScope ps__(ctx, st, K_UF_ACTIVATE);
                    uf_union_list_kernel<<<ctx->sm_count * 16, 256, 0, st>>>(w.parent, acls, folded, k - 1, k, 6, (int)nz, (int)ny,
                                                                             (int)nx, w.uf_list, nullptr, jtime, start);
                }
                LAUNCH_CHECK(ctx);
            }
        }
        {
            ProfScope ps__(ctx, st, K_UF_MARK);
            uf_resolve_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(
                w.parent, w.cls, jtime, w.rcls, n, kmin);
        }
        LAUNCH_CHECK(ctx);
    }
    const uint8_t *cmap = al ? w.rcls : w.cls;
    const int *gate = al ? w.gate : nullptr;
    const bool bit_ok = ctx->algo == PSB200_ALGO_FAST && ctx->bit_tmax > 0 && (nx % 32 == 0) && nz <= 65535 && n < (1LL << 35) &&
                        ((((uintptr_t)idx | (uintptr_t)cmap) & 15u) == 0);
    bool wmask_ready = false;
    int packed_lo = 0, packed_hi = 0;
    for (int k = 0; k < nT; ++k) {
        const uint32_t T = T_host[k];
        if (al) {
            uf_gate_kernel<<<1, 1, 0, st>>>(w.gate, w.gate + 1, k);
            LAUNCH_CHECK(ctx);
        }
        if (bit_ok && T <= (uint32_t)ctx->bit_tmax) {
            // small radii: bit-parallel ball dilation (thresholds descend: every later radius too)
            if (!wmask_ready) {
                const int64_t nwords = n / 32;
                if (k == 0 && !(flags & PSB200_FLAG_IDX_PREINIT))
                    CUDA_TRY(cudaMemsetAsync(w.written, 0, (size_t)nwords * 4, st));
                else {
                    rc = lt_wmask_impl(ctx, idx, w.written, nwords, st);
                    if (rc) return rc;
                }
                wmask_ready = true;
            }
            rc = lt_bit_step(ctx, w, cmap, idx, k, nT, T, nz, ny, nx, gate, st, packed_lo, packed_hi);
            if (rc) return rc;
            continue;
        }
        if (ctx->algo == PSB200_ALGO_GENERIC) {
            // reference-shaped path: full EDT of ~seeds, then fill <=> d2' < T  (F:1191)
            rc = launch_x<1>(ctx, cmap, w.gen_d2, nz * ny, (int)nx, k, st);
            if (rc) return rc;
            const EdtStoreFill fillst{idx, T, (uint8_t)(k + 1)};
            if (nz > 1) {
                if (ny > 1) {
                    rc = launch_col(ctx, 1, w.gen_d2, EdtStoreU32{w.gen_d2}, nz, ny, nx, w.gen_stk, w.gen_stk_bytes, st, K_GEN_Y);
                    if (rc) return rc;
                }
                rc = launch_col(ctx, 0, w.gen_d2, fillst, nz, ny, nx, w.gen_stk, w.gen_stk_bytes, st, K_GEN_Z);
            } else if (ny > 1) {
                rc = launch_col(ctx, 1, w.gen_d2, fillst, nz, ny, nx, w.gen_stk, w.gen_stk_bytes, st, K_GEN_Y);
            } else {
                {
                    ProfScope ps__(ctx, st, K_GEN_Z);
                    fill_from_d2_kernel<<<g, 256, 0, st>>>(w.gen_d2, fillst, n);
                }
                LAUNCH_CHECK(ctx);
            }
            if (rc) return rc;
            continue;
        }
        if (T == 1) {
            {
                ProfScope ps__(ctx, st, K_LT_POINT);
                lt_point_kernel<<<g, 256, 0, st>>>(cmap, idx, n, k, (uint32_t)(k + 1), gate);
            }
            LAUNCH_CHECK(ctx);
            continue;
        }
        if (streaming_ok(ny, nx, T, cmap, w.reach, w.gx) && (((uintptr_t)idx & 15u) == 0)) {
            const uint32_t *bits = nullptr;
            if (bit_ok && ctx->xbits) {
                rc = lt_seed_bits(ctx, w, cmap, k, nT, n / 32, st, packed_lo, packed_hi, &bits);
                if (rc) return rc;
            }
            rc = lt_xy_stream_impl(ctx, cmap, k, T, w.gx, w.reach, nz, ny, nx, gate, st, bits, w.xflag);
            if (rc) return rc;
            rc = lt_z_stream_impl(ctx, w.reach, nullptr, 0, nullptr, 0, idx, k, T, nz, ny, nx, gate, st);
        } else {
            rc = lt_xy_impl(ctx, cmap, k, T, w.reach, nz, ny, nx, gate, st);
            if (rc) return rc;
            rc = lt_z_impl(ctx, w.reach, nullptr, 0, nullptr, 0, idx, k, T, nz, ny, nx, gate, st);
        }
        if (rc) return rc;
    }
    return PSB200_OK;
}

extern "C" int psb200_expand_idx_f64(psb200_ctx *ctx, const uint8_t *idx, const double *lut_host, int nlut,
                                     double *out, int64_t n, int flags, psb200_stream stream)
{
    if (!ctx || !idx || !lut_host || !out || nlut < 1 || nlut > 254 || n < 0)
        return fail(PSB200_ERR_INVALID, "expand_idx_f64: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    LutArg lut;
    memset(&lut, 0, sizeof(lut));
    for (int i = 0; i < nlut; ++i) lut.v[i] = lut_host[i];
    {
        ProfScope ps__(ctx, st, K_EXPAND);
        lt_expand_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(
            idx, lut, nlut, out, n, (flags & PSB200_FLAG_EXPAND_MERGE) ? 1 : 0);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ---- host-result epilogue (host_epilogue.cuh)
#include "host_epilogue.cuh"
static HostEpilogueStreams g_epilogue_streams[64];         // per device, created on first use

static int launch_expand_chunk(psb200_ctx *ctx, const uint8_t *idx, const double *lut_host, int nlut, double *out,
                               int64_t n, cudaStream_t st)
{
    return psb200_expand_idx_f64(ctx, idx, lut_host, nlut, out, n, 0, (psb200_stream)st);
}

extern "C" int psb200_expand_idx_f64_to_host(psb200_ctx *ctx, const uint8_t *idx, const double *lut_host, int nlut,
                                             double *out_host, int64_t n, uint8_t *stage_host, size_t stage_bytes,
                                             void *ws, size_t ws_bytes, int cpu_permille, int nthreads,
                                             int flags, psb200_stream stream)
{
    if (!ctx || !idx || !lut_host || !out_host || nlut < 1 || nlut > 254 || n < 0)
        return fail(PSB200_ERR_INVALID, "expand_idx_f64_to_host: bad argument");
    if (cpu_permille < 0 || cpu_permille > 1000 || nthreads < 0 || nthreads > 1024)
        return fail(PSB200_ERR_INVALID, "expand_idx_f64_to_host: cpu_permille must be in [0,1000], nthreads in [0,1024]");
    if (ctx->device < 0 || ctx->device >= 64) return fail(PSB200_ERR_UNSUPPORTED, "expand_idx_f64_to_host: device index");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    (void)flags;
    if (nthreads == 0) nthreads = (int)std::thread::hardware_concurrency();
    return host_epilogue_run(ctx, g_epilogue_streams[ctx->device], idx, lut_host, nlut, out_host, n, stage_host,
                             stage_bytes, ws, ws_bytes, cpu_permille, nthreads, (cudaStream_t)stream,
                             launch_expand_chunk);
}

extern "C" int psb200_upload_mask_u8(psb200_ctx *ctx, const uint8_t *src_host, int64_t n, uint8_t *dst,
                                     uint8_t *stage_host, size_t stage_bytes, void *ws, size_t ws_bytes,
                                     int nthreads, psb200_stream stream)
{
    if (!ctx || !src_host || !dst || !stage_host || !ws || n < 0 || nthreads < 0 || nthreads > 1024)
        return fail(PSB200_ERR_INVALID, "upload_mask_u8: bad argument");
    const size_t nb = (size_t)((n + 7) >> 3);
    if (stage_bytes < nb || ws_bytes < nb)
        return fail(PSB200_ERR_WORKSPACE, "upload_mask_u8 needs %zu bytes of staging and of device workspace", nb);
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (nthreads == 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    return host_upload_mask(ctx, src_host, n, dst, stage_host, reinterpret_cast<uint8_t *>(ws), nthreads,
                            (cudaStream_t)stream);
}

extern "C" int psb200_mark_written(psb200_ctx *ctx, const double *out, uint8_t *idx, int64_t n,
                                   psb200_stream stream)
{
    if (!ctx || !out || !idx || n < 0) return fail(PSB200_ERR_INVALID, "mark_written: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        ProfScope ps__(ctx, (cudaStream_t)stream, K_MARK_WRITTEN);
        lt_mark_written_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, (cudaStream_t)stream>>>(out, idx, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ---- step-level access-limited flooding for z-slab shards (flood halos)
static int uf_args(const char *who, psb200_ctx *ctx, const void *parent, const void *cls, const uint8_t *inlets,
                   int inlet_mode, int ndim, int64_t nz, int64_t ny, int64_t nx, int64_t z0, int64_t nz_global)
{
    if (!ctx || !parent || !cls) return fail(PSB200_ERR_INVALID, "%s: NULL argument", who);
    if (inlet_mode != PSB200_INLETS_FACES && inlet_mode != PSB200_INLETS_MASK)
        return fail(PSB200_ERR_INVALID, "%s: inlet_mode must be FACES or MASK", who);
    if (inlet_mode == PSB200_INLETS_MASK && !inlets) return fail(PSB200_ERR_INVALID, "%s: inlet mask is NULL", who);
    if (ndim < 1 || ndim > 3) return fail(PSB200_ERR_INVALID, "%s: ndim must be 1..3", who);
    int rc = check_dims(who, nz, ny, nx);
    if (rc) return rc;
    if (z0 < 0 || z0 + nz > nz_global || nz_global > PSB200_MAX_DIM)
        return fail(PSB200_ERR_INVALID, "%s: slab [z0, z0+nz) outside [0, nz_global)", who);
    if (nz * ny * nx > 0xFFFFFFF0LL) return fail(PSB200_ERR_UNSUPPORTED, "%s: flooding supports < 2^32 voxels per GPU", who);
    return PSB200_OK;
}

extern "C" int psb200_uf_begin(psb200_ctx *ctx, const uint8_t *cls, uint8_t *rcls, uint32_t *parent,
                               const uint8_t *inlets, int inlet_mode, int ndim, int64_t nz, int64_t ny,
                               int64_t nx, int64_t z0, int64_t nz_global, psb200_stream stream)
{
    int rc = uf_args("uf_begin", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    if (!rcls) return fail(PSB200_ERR_INVALID, "uf_begin: rcls is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = nz * ny * nx;
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    {
        ProfScope ps__(ctx, st, K_UF_INIT);
        uf_rcls_init_kernel<<<g, 256, 0, st>>>(cls, rcls, n);
    }
    LAUNCH_CHECK(ctx);
    {
        ProfScope ps__(ctx, st, K_UF_INIT);
        uf_init_kernel<<<g, 256, 0, st>>>(parent, inl, (int)nz, (int)ny, (int)nx, nullptr);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" size_t psb200_uf_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx)
{
    if (!ctx) return 0;
    return (uf_list_entries(nz * ny * nx) + 64) * sizeof(uint32_t) + 512;
}

extern "C" int psb200_uf_activate(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                                  int inlet_mode, int ndim, int klo, int khi, int64_t nz, int64_t ny,
                                  int64_t nx, int64_t z0, int64_t nz_global, void *ws, size_t ws_bytes,
                                  psb200_stream stream)
{
    int rc = uf_args("uf_activate", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    const size_t need = (uf_list_entries(nz * ny * nx) + 64) * sizeof(uint32_t);
    if (!base || ws_bytes < need + (size_t)(base - (char *)ws))
        return fail(PSB200_ERR_WORKSPACE, "uf_activate needs %zu workspace bytes, got %zu", need + 256, ws_bytes);
    CUDA_TRY(cudaSetDevice(ctx->device));
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    return uf_activate_impl(ctx, parent, cls, inl, klo, khi, 6, nz, ny, nx, reinterpret_cast<uint32_t *>(base),
                            (cudaStream_t)stream);
}

// Record store of the step-level flood (psb200_uf_*_records): the same row-rooted forest + link records + join
// times as the single-GPU loop (flood_kernels.cuh), built once for all radius indices and kept by the caller across
// the radius loop.
static size_t carve_uf_records(char *base, int64_t nz, int64_t ny, int64_t nx, LtWorkspace &w, uint8_t **acls,
                               uint8_t **jtime)
{
    const size_t n = (size_t)nz * ny * nx;
    Carver c{base, 0};
    w.gate = c.take<int>(64);
    w.uf_bins = c.take<uint32_t>(3 * (UF_NTIMES * UF_MAXFAM + 64));
    *acls = c.take<uint8_t>(n + 16);
    *jtime = c.take<uint8_t>(n + 16);
    const int64_t segs = nz * ny * ((nx + UF_SEGX - 1) / UF_SEGX);
    w.uf_chunks = (segs + UF_CHUNK - 1) / UF_CHUNK;
    w.uf_cap = (uint32_t)(2 * UF_CHUNK * (nx < UF_SEGX ? nx : UF_SEGX) + 2 * UF_CHUNK + 64);
    w.uf_list_cap = (size_t)w.uf_chunks * w.uf_cap;
    w.uf_list = c.take<uint32_t>(w.uf_list_cap);
    w.uf_raw = c.take<uint32_t>((size_t)w.uf_chunks * w.uf_cap);
    w.uf_ccount = c.take<uint32_t>((size_t)w.uf_chunks + 64);
    return c.off + 256;
}

struct UfRecords {
    LtWorkspace w;
    uint8_t *acls, *jtime;
    const uint32_t *start;
};

static int uf_records_open(const char *who, void *rec, size_t rec_bytes, int64_t nz, int64_t ny, int64_t nx, UfRecords &r)
{
    if (!rec) return fail(PSB200_ERR_INVALID, "%s: record store is NULL", who);
    char *base = (char *)(((uintptr_t)rec + 255) & ~(uintptr_t)255);
    r.w = LtWorkspace{};
    const size_t need = carve_uf_records(base, nz, ny, nx, r.w, &r.acls, &r.jtime);
    if (rec_bytes < need + (size_t)(base - (char *)rec))
        return fail(PSB200_ERR_WORKSPACE, "%s needs %zu bytes of record store, got %zu", who, need + 256, rec_bytes);
    r.start = r.w.uf_bins + UF_NTIMES * 6 + 64;          // (uf_forest_impl: 6-connectivity, 6 slices per radius index)
    return PSB200_OK;
}

extern "C" size_t psb200_uf_records_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx)
{
    if (!ctx || nz < 1 || ny < 1 || nx < 1) return 0;
    LtWorkspace w{};
    uint8_t *acls = nullptr, *jtime = nullptr;
    return carve_uf_records(nullptr, nz, ny, nx, w, &acls, &jtime) + 256;
}

extern "C" int psb200_uf_begin_records(psb200_ctx *ctx, const uint8_t *cls, uint32_t *parent, const uint8_t *inlets,
                                       int inlet_mode, int ndim, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                                       int64_t nz_global, void *rec, size_t rec_bytes, psb200_stream stream)
{
    int rc = uf_args("uf_begin_records", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    UfRecords r;
    rc = uf_records_open("uf_begin_records", rec, rec_bytes, nz, ny, nx, r);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    r.w.parent = parent;
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    const uint32_t *start = nullptr;
    bool fits = true;
    return uf_forest_impl(ctx, r.w, inl, cls, r.acls, r.jtime, 6, nz, ny, nx, (cudaStream_t)stream, &start, &fits);
}

extern "C" int psb200_uf_activate_records(psb200_ctx *ctx, uint32_t *parent, int klo, int khi, int64_t nz, int64_t ny,
                                          int64_t nx, void *rec, size_t rec_bytes, psb200_stream stream)
{
    if (!ctx || !parent) return fail(PSB200_ERR_INVALID, "uf_activate_records: NULL argument");
    int rc = check_dims("uf_activate_records", nz, ny, nx);
    if (rc) return rc;
    if (klo < -1 || khi >= UF_NTIMES || klo > khi) return fail(PSB200_ERR_INVALID, "uf_activate_records: bad radius range");
    UfRecords r;
    rc = uf_records_open("uf_activate_records", rec, rec_bytes, nz, ny, nx, r);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    r.w.parent = parent;
    for (int k = klo + 1; k <= khi; ++k) {
        rc = uf_union_records(ctx, r.w, r.start, k, 6, ny, nx, r.jtime, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return PSB200_OK;
}

extern "C" int psb200_uf_face_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                                      int inlet_mode, int ndim, int k, int64_t zplane, uint8_t *flags_out, int64_t nz,
                                      int64_t ny, int64_t nx, int64_t z0, int64_t nz_global, void *rec, size_t rec_bytes,
                                      psb200_stream stream)
{
    int rc = uf_args("uf_face_records", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    if (!flags_out || zplane < 0 || zplane >= nz) return fail(PSB200_ERR_INVALID, "uf_face_records: bad plane / output");
    UfRecords r;
    rc = uf_records_open("uf_face_records", rec, rec_bytes, nz, ny, nx, r);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    {
        ProfScope ps__(ctx, st, K_UF_FACE);
        uf_face_kernel<<<grid_for(ny * nx, 256, ctx->sm_count, 8), 256, 0, st>>>(parent, cls, inl, k, (int)zplane, (int)nz,
                                                                               (int)ny, (int)nx, flags_out, r.jtime);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_uf_inject_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                                        int inlet_mode, int ndim, int k, int64_t zplane, const uint8_t *nb_flags,
                                        int *changed_dev, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                                        int64_t nz_global, void *rec, size_t rec_bytes, psb200_stream stream)
{
    int rc = uf_args("uf_inject_records", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    if (!nb_flags || !changed_dev || zplane < 0 || zplane >= nz)
        return fail(PSB200_ERR_INVALID, "uf_inject_records: bad plane / flags / changed pointer");
    UfRecords r;
    rc = uf_records_open("uf_inject_records", rec, rec_bytes, nz, ny, nx, r);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    {
        ProfScope ps__(ctx, st, K_UF_FACE);
        uf_inject_kernel<<<grid_for(ny * nx, 256, ctx->sm_count, 8), 256, 0, st>>>(parent, cls, inl, k, (int)zplane, (int)nz,
                                                                                 (int)ny, (int)nx, nb_flags, changed_dev,
                                                                                 r.jtime);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_uf_resolve_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, uint8_t *rcls, int64_t nz,
                                         int64_t ny, int64_t nx, void *rec, size_t rec_bytes, psb200_stream stream)
{
    if (!ctx || !parent || !cls || !rcls) return fail(PSB200_ERR_INVALID, "uf_resolve_records: NULL argument");
    int rc = check_dims("uf_resolve_records", nz, ny, nx);
    if (rc) return rc;
    UfRecords r;
    rc = uf_records_open("uf_resolve_records", rec, rec_bytes, nz, ny, nx, r);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = nz * ny * nx;
    int *kmin = r.w.gate + 1;
    CUDA_TRY(cudaMemsetAsync(kmin, 0x7F, sizeof(int), st));
    {
        ProfScope ps__(ctx, st, K_UF_MARK);
        uf_compress_kernel<<<ctx->sm_count * 6, 256, 0, st>>>(parent, r.acls, (int)nz, (int)ny, (int)nx);
    }
    LAUNCH_CHECK(ctx);
    {
        ProfScope ps__(ctx, st, K_UF_MARK);
        uf_resolve_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(parent, cls, r.jtime, rcls, n, kmin);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_uf_face(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                              int inlet_mode, int ndim, int k, int64_t zplane, uint8_t *flags_out, int64_t nz,
                              int64_t ny, int64_t nx, int64_t z0, int64_t nz_global, psb200_stream stream)
{
    int rc = uf_args("uf_face", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    if (!flags_out || zplane < 0 || zplane >= nz) return fail(PSB200_ERR_INVALID, "uf_face: bad plane / output");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    {
        ProfScope ps__(ctx, st, K_UF_FACE);
        uf_face_kernel<<<grid_for(ny * nx, 256, ctx->sm_count, 16), 256, 0, st>>>(
            parent, cls, inl, k, (int)zplane, (int)nz, (int)ny, (int)nx, flags_out);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_uf_inject(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                                int inlet_mode, int ndim, int k, int64_t zplane, const uint8_t *nb_flags,
                                int *changed_dev, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                                int64_t nz_global, psb200_stream stream)
{
    int rc = uf_args("uf_inject", ctx, parent, cls, inlets, inlet_mode, ndim, nz, ny, nx, z0, nz_global);
    if (rc) return rc;
    if (!nb_flags || !changed_dev || zplane < 0 || zplane >= nz)
        return fail(PSB200_ERR_INVALID, "uf_inject: bad plane / flags / changed pointer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    InletSpec inl{inlet_mode, ndim, inlets, (int)z0, (int)nz_global};
    {
        ProfScope ps__(ctx, st, K_UF_FACE);
        uf_inject_kernel<<<grid_for(ny * nx, 256, ctx->sm_count, 16), 256, 0, st>>>(
            parent, cls, inl, k, (int)zplane, (int)nz, (int)ny, (int)nx, nb_flags, changed_dev);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_uf_mark(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, uint8_t *rcls, int k,
                              int *any_dev, int64_t n, psb200_stream stream)
{
    if (!ctx || !parent || !cls || !rcls || !any_dev || n < 0) return fail(PSB200_ERR_INVALID, "uf_mark: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps__(ctx, st, K_UF_MARK);
        uf_mark_kernel<<<grid_for((n + 15) / 16, 256, ctx->sm_count, 16), 256, 0, st>>>(parent, cls, rcls, k, n, any_dev);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// -------------------------------------------------------------------------------- flood
extern "C" int psb200_flood(psb200_ctx *ctx, const uint8_t *mask, const uint8_t *inlets, uint8_t *out,
                            int conn, int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                            psb200_stream stream)
{
    if (!ctx || !mask || !inlets || !out) return fail(PSB200_ERR_INVALID, "flood: NULL argument");
    int rc = check_dims("flood", nz, ny, nx);
    if (rc) return rc;
    int c3;
    if (conn == 6 || conn == 4) c3 = 6;
    else if (conn == 26 || conn == 8) c3 = 26;
    else return fail(PSB200_ERR_INVALID, "flood: conn must be 4/8 (2-D) or 6/26 (3-D), got %d", conn);
    if ((conn == 4 || conn == 8) && nz != 1)
        return fail(PSB200_ERR_INVALID, "flood: conn %d is 2-D connectivity but nz=%lld", conn, (long long)nz);
    const int64_t n = nz * ny * nx;
    if (n > 0xFFFFFFF0LL) return fail(PSB200_ERR_UNSUPPORTED, "flood supports < 2^32 voxels per GPU");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    psb200_ctx tmp = *ctx;
    tmp.algo = PSB200_ALGO_FAST;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    LtWorkspace w = carve_lt(&tmp, base, nz, ny, nx, PSB200_INLETS_MASK);
    if (!ws || w.total + 256 > ws_bytes)
        return fail(PSB200_ERR_WORKSPACE, "flood needs %zu workspace bytes, got %zu", w.total + 256, ws_bytes);
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    InletSpec inl{PSB200_INLETS_MASK, 3, inlets, 0, (int)nz};
    CUDA_TRY(cudaMemsetAsync(w.gate, 0, sizeof(int), st));
    {
        ProfScope ps__(ctx, st, K_FLOOD_MISC);
        flood_cls_kernel<<<g, 256, 0, st>>>(mask, w.cls, n);
    }
    LAUNCH_CHECK(ctx);
    bool done = false;
    if (ctx->uf_records) {
        // row-rooted forest + link records (the activation map goes to the reach buffer, unused here)
        const uint32_t *start = nullptr;
        rc = uf_forest_impl(ctx, w, inl, w.cls, w.reach, nullptr, c3, nz, ny, nx, st, &start, &done);
        if (rc) return rc;
        if (done) {
            rc = uf_union_records(ctx, w, start, 0, c3, ny, nx, nullptr, st);
            if (rc) return rc;
            {
                ProfScope ps__(ctx, st, K_UF_MARK);
                uf_compress_kernel<<<ctx->sm_count * 6, 256, 0, st>>>(w.parent, w.reach, (int)nz, (int)ny, (int)nx);
            }
            LAUNCH_CHECK(ctx);
            {
                ProfScope ps__(ctx, st, K_UF_MARK);
                uf_reach_out_kernel<<<g, 256, 0, st>>>(w.parent, w.cls, out, n);
            }
            LAUNCH_CHECK(ctx);
            return PSB200_OK;
        }
    }
    if (!done) {
        {
            ProfScope ps__(ctx, st, K_UF_INIT);
            uf_rcls_init_kernel<<<g, 256, 0, st>>>(w.cls, w.rcls, n);
        }
        LAUNCH_CHECK(ctx);
        {
            ProfScope ps__(ctx, st, K_UF_INIT);
            uf_init_kernel<<<g, 256, 0, st>>>(w.parent, inl, (int)nz, (int)ny, (int)nx, nullptr);
        }
        LAUNCH_CHECK(ctx);
        rc = uf_step(ctx, w, inl, -1, 0, c3, nz, ny, nx, st);
        if (rc) return rc;
    }
    {
        ProfScope ps__(ctx, st, K_FLOOD_MISC);
        flood_out_kernel<<<g, 256, 0, st>>>(w.rcls, out, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// Flood of nested sets: cls[v] = first step at which voxel v is a node (254: never, 255: not a node), rcls[v] = first
// step at which it is a node connected to the inlets.  inlets_in_set = 0: inlet voxels are nodes from step 0 on
// (trim_disconnected_blobs, F:1265); 1: an inlet voxel only counts from the step at which it joins the set
// (find_trapped_regions labels `seq >= i` and looks up the outlets inside it, F:131-137).  The machinery of
// the access-limited radius loop (join times) for any caller whose per-step masks are nested: drainage's pressure
// steps (simulations/_drainage.py:133-154).
extern "C" int psb200_flood_classes(psb200_ctx *ctx, const uint8_t *cls, const uint8_t *inlets, int inlets_in_set,
                                    uint8_t *rcls, int nsteps, int conn, int64_t nz, int64_t ny, int64_t nx, void *ws,
                                    size_t ws_bytes, psb200_stream stream)
{
    if (!ctx || !cls || !inlets || !rcls) return fail(PSB200_ERR_INVALID, "flood_classes: NULL argument");
    int rc = check_dims("flood_classes", nz, ny, nx);
    if (rc) return rc;
    if (nsteps < 1 || nsteps > UF_NTIMES) return fail(PSB200_ERR_INVALID, "flood_classes: 1..254 steps, got %d", nsteps);
    int c3;
    if (conn == 6 || conn == 4) c3 = 6;
    else if (conn == 26 || conn == 8) c3 = 26;
    else return fail(PSB200_ERR_INVALID, "flood_classes: conn must be 4/8 (2-D) or 6/26 (3-D), got %d", conn);
    if ((conn == 4 || conn == 8) && nz != 1)
        return fail(PSB200_ERR_INVALID, "flood_classes: conn %d is 2-D connectivity but nz=%lld", conn, (long long)nz);
    const int64_t n = nz * ny * nx;
    if (n > 0xFFFFFFF0LL) return fail(PSB200_ERR_UNSUPPORTED, "flood_classes supports < 2^32 voxels per GPU");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    psb200_ctx tmp = *ctx;
    tmp.algo = PSB200_ALGO_FAST;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    LtWorkspace w = carve_lt(&tmp, base, nz, ny, nx, PSB200_INLETS_MASK);
    if (!ws || w.total + 256 > ws_bytes)
        return fail(PSB200_ERR_WORKSPACE, "flood_classes needs %zu workspace bytes, got %zu", w.total + 256, ws_bytes);
    InletSpec inl{inlets_in_set ? 4 : PSB200_INLETS_MASK, 3, inlets, 0, (int)nz};
    uint8_t *jtime = w.reach, *acls = w.rcls;
    int *kmin = w.gate + 1;
    CUDA_TRY(cudaMemsetAsync(kmin, 0x7F, sizeof(int), st));
    const uint32_t *start = nullptr;
    bool fits = false;
    if (ctx->uf_records) {
        rc = uf_forest_impl(ctx, w, inl, cls, acls, jtime, c3, nz, ny, nx, st, &start, &fits);
        if (rc) return rc;
    }
    if (fits) {
        for (int k = 0; k < nsteps; ++k) {
            rc = uf_union_records(ctx, w, start, k, c3, ny, nx, jtime, st);
            if (rc) return rc;
        }
        {
            ProfScope ps__(ctx, st, K_UF_MARK);
            uf_compress_kernel<<<ctx->sm_count * 6, 256, 0, st>>>(w.parent, acls, (int)nz, (int)ny, (int)nx);
        }
        LAUNCH_CHECK(ctx);
    } else {
        // per-voxel job lists (record overflow with 26-connectivity, or the option switched off)
        {
            ProfScope ps__(ctx, st, K_UF_INIT);
            uf_init_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(w.parent, inl, (int)nz, (int)ny, (int)nx, jtime, cls, acls);
        }
        LAUNCH_CHECK(ctx);
        const InletSpec folded{3, 3, nullptr, 0, (int)nz};
        for (int k = 0; k < nsteps; ++k) {
            rc = uf_activate_impl(ctx, w.parent, acls, folded, k - 1, k, c3, nz, ny, nx, w.uf_list, st, jtime);
            if (rc) return rc;
        }
    }
    {
        ProfScope ps__(ctx, st, K_UF_MARK);
        uf_resolve_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(w.parent, cls, jtime, rcls, n, kmin);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ------------------------------------------------------------------------- blobs generator
// (blobs_kernels.cuh; reference generators/_imgen.py:1023-1051, tools/_funcs.py:963-969)
extern "C" int psb200_noise_philox_f64(psb200_ctx *ctx, double *out, int64_t n, uint64_t seed, uint64_t first,
                                       psb200_stream stream)
{
    if (!ctx || !out || n < 0) return fail(PSB200_ERR_INVALID, "noise_philox_f64: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps__(ctx, st, K_BLOBS);
        noise_philox_kernel<<<grid_for(n / 2 + 1, 256, ctx->sm_count, 16), 256, 0, st>>>(out, n, seed, first);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" size_t psb200_gauss_workspace_bytes(const psb200_ctx *ctx, int radius)
{
    if (!ctx || radius < 0) return 0;
    return (size_t)(radius + 1) * sizeof(double) + 256;
}

extern "C" int psb200_gauss_axis_f64(psb200_ctx *ctx, const double *in, double *out, int axis,
                                     const double *w_host, int radius, int64_t nz, int64_t ny, int64_t nx,
                                     int64_t z_out0, int64_t z_in0, int64_t nz_in, int64_t nz_glob, void *ws,
                                     size_t ws_bytes, psb200_stream stream)
{
    if (!ctx || !in || !out || in == out || !w_host || radius < 0 || axis < 0 || axis > 2)
        return fail(PSB200_ERR_INVALID, "gauss_axis_f64: bad argument");
    int rc = check_dims("gauss_axis_f64", nz, ny, nx);
    if (rc) return rc;
    if (axis == 0 && (nz_glob < 1 || z_out0 < 0 || z_out0 + nz > nz_glob || z_in0 < 0 || z_in0 + nz_in > nz_glob))
        return fail(PSB200_ERR_INVALID, "gauss_axis_f64: slab outside the global volume");
    if (axis == 0) {
        // every (reflected) tap of every output plane must lie inside the input planes
        int64_t need_lo = nz_glob, need_hi = -1;
        for (int64_t i = z_out0 - radius; i <= z_out0 + nz - 1 + radius; ++i) {
            const int64_t per = 2 * nz_glob;
            int64_t m = i % per;
            if (m < 0) m += per;
            if (m >= nz_glob) m = per - 1 - m;
            if (m < need_lo) need_lo = m;
            if (m > need_hi) need_hi = m;
        }
        if (need_lo < z_in0 || need_hi >= z_in0 + nz_in)
            return fail(PSB200_ERR_INVALID, "gauss_axis_f64: input planes [%lld,%lld) do not cover the filter reach "
                        "[%lld,%lld] of output planes [%lld,%lld)", (long long)z_in0, (long long)(z_in0 + nz_in),
                        (long long)need_lo, (long long)need_hi, (long long)z_out0, (long long)(z_out0 + nz));
    }
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    const size_t wbytes = (size_t)(radius + 1) * sizeof(double);
    if (!base || ws_bytes < wbytes + (size_t)(base - (char *)ws))
        return fail(PSB200_ERR_WORKSPACE, "gauss_axis_f64 needs %zu workspace bytes", wbytes + 256);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    double *wdev = reinterpret_cast<double *>(base);
    CUDA_TRY(cudaMemcpyAsync(wdev, w_host, wbytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));                 // w_host may be a temporary of the caller
    const int64_t plane = ny * nx;
    if (axis == 2) {
        const size_t smem = (size_t)(GX_SEG + 3 * radius + 1) * sizeof(double);
        if ((int)smem > ctx->max_smem_optin) return fail(PSB200_ERR_UNSUPPORTED, "gauss_axis_f64: radius %d too large", radius);
        CUDA_TRY(cudaFuncSetAttribute(gauss_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t jobs = nz * ny * ((nx + GX_SEG - 1) / GX_SEG);
        {
            ProfScope ps__(ctx, st, K_BLOBS);
            gauss_x_kernel<<<grid_for(jobs, 1, ctx->sm_count, 8), 256, smem, st>>>(in, out, nz * ny, (int)nx, radius, wdev);
        }
        LAUNCH_CHECK(ctx);
        return PSB200_OK;
    }
    int rows = 128;
    while (rows > 8 && (size_t)(rows + 2 * radius) * 256 + wbytes + 8 > (size_t)ctx->max_smem_optin / 2) rows >>= 1;
    const size_t smem = (size_t)(rows + 2 * radius) * 256 + wbytes + 8;
    if ((int)smem > ctx->max_smem_optin) return fail(PSB200_ERR_UNSUPPORTED, "gauss_axis_f64: radius %d too large", radius);
    CUDA_TRY(cudaFuncSetAttribute(gauss_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ncols, stride, outer, os_in, os_out, g0_out, g0_in, n_glob;
    int n_out, n_in;
    if (axis == 1) { ncols = nx; stride = nx; n_out = n_in = (int)ny; g0_out = g0_in = 0; n_glob = ny; outer = nz; os_in = os_out = plane; }
    else { ncols = plane; stride = plane; n_out = (int)nz; n_in = (int)nz_in; g0_out = z_out0; g0_in = z_in0; n_glob = nz_glob; outer = 1; os_in = os_out = 0; }
    const int64_t jobs = ((ncols + 31) / 32) * ((n_out + rows - 1) / rows) * outer;
    {
        ProfScope ps__(ctx, st, K_BLOBS);
        gauss_col_kernel<<<grid_for(jobs, 1, ctx->sm_count, 4), 256, smem, st>>>(in, out, ncols, stride, n_out, g0_out, n_in, g0_in,
                                                                             n_glob, outer, os_in, os_out, radius, rows, wdev);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_stats_chunks(void) { return ST_CHUNKS; }

extern "C" int psb200_stats_f64(psb200_ctx *ctx, const double *x, int64_t nplanes, int64_t plane, double mean, int mode,
                                double *part, psb200_stream stream)
{
    if (!ctx || !x || !part || nplanes < 1 || plane < 1 || mode < 0 || mode > 3)
        return fail(PSB200_ERR_INVALID, "stats_f64: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps__(ctx, st, K_BLOBS);
        stats_kernel<<<grid_for(nplanes * ST_CHUNKS, 1, ctx->sm_count, 8), 256, 0, st>>>(x, nplanes, plane, mean, mode, part);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_blobs_finish(psb200_ctx *ctx, const double *f, int64_t n, double mean, double sd, double fmin_,
                                   double fmax_, double porosity, uint8_t *out_u8, double *out_f64, psb200_stream stream)
{
    if (!ctx || !f || n < 0 || (!out_u8 && !out_f64)) return fail(PSB200_ERR_INVALID, "blobs_finish: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps__(ctx, st, K_BLOBS);
        blobs_finish_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(f, n, mean, sd, fmin_, fmax_, porosity,
                                                                              out_u8, out_u8 ? nullptr : out_f64);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ------------------------------------------------- radius-map post-processing on the index form
// (sizemap_kernels.cuh; reference filters/_size_seq_satn.py:16-221, metrics/_funcs.py:558-632, 1073-1090)
extern "C" int psb200_hist_idx(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask, int64_t n, int K,
                               uint64_t *counts, psb200_stream stream)
{
    if (!ctx || !idx || !counts || n < 0 || K < 1 || K > 65536 || (idx_bytes != 1 && idx_bytes != 2) || (idx_bytes == 1 && K > 256))
        return fail(PSB200_ERR_INVALID, "hist_idx: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)(mask ? 2 * K : K) * sizeof(uint64_t), st));
    if (n == 0) return PSB200_OK;
    const int g = grid_for(n, 256, ctx->sm_count, 8);
    {
        ProfScope ps__(ctx, st, K_SIZEMAP);
        if (idx_bytes == 1)
            hist_idx_kernel<uint8_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(idx), mask, n, K,
                                                      reinterpret_cast<unsigned long long *>(counts));
        else
            hist_idx_kernel<uint16_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint16_t *>(idx), mask, n, K,
                                                       reinterpret_cast<unsigned long long *>(counts));
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_expand_lut8(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask,
                                  const uint64_t *lut, void *out, int64_t n, int K, psb200_stream stream)
{
    if (!ctx || !idx || !lut || !out || n < 0 || K < 1 || K > 65536 || (idx_bytes != 1 && idx_bytes != 2))
        return fail(PSB200_ERR_INVALID, "expand_lut8: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    {
        ProfScope ps__(ctx, st, K_SIZEMAP);
        if (idx_bytes == 1)
            expand_lut8_kernel<uint8_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(idx), mask, lut,
                                                         reinterpret_cast<uint64_t *>(out), n, K);
        else
            expand_lut8_kernel<uint16_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint16_t *>(idx), mask, lut,
                                                          reinterpret_cast<uint64_t *>(out), n, K);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_expand_lut1(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask,
                                  const uint8_t *lut, uint8_t *out, int64_t n, int K, psb200_stream stream)
{
    if (!ctx || !idx || !lut || !out || n < 0 || K < 1 || K > 65536 || (idx_bytes != 1 && idx_bytes != 2))
        return fail(PSB200_ERR_INVALID, "expand_lut1: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    {
        ProfScope ps__(ctx, st, K_SIZEMAP);
        if (idx_bytes == 1)
            expand_lut1_kernel<uint8_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(idx), mask, lut, out, n, K);
        else
            expand_lut1_kernel<uint16_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint16_t *>(idx), mask, lut, out, n, K);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_distinct64(psb200_ctx *ctx, const uint64_t *x, int64_t n, uint64_t *table, uint32_t cap,
                                 int *overflow, psb200_stream stream)
{
    if (!ctx || !x || !table || !overflow || n < 0 || cap < 2 || (cap & (cap - 1)))
        return fail(PSB200_ERR_INVALID, "distinct64: bad argument (cap must be a power of two)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    fill_u64_kernel<<<grid_for(cap, 256, ctx->sm_count, 4), 256, 0, st>>>(reinterpret_cast<unsigned long long *>(table), cap,
                                                                         DISTINCT_EMPTY);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemsetAsync(overflow, 0, sizeof(int), st));
    if (n == 0) return PSB200_OK;
    {
        ProfScope ps__(ctx, st, K_SIZEMAP);
        distinct64_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, st>>>(
            x, n, reinterpret_cast<unsigned long long *>(table), cap, overflow);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_index_of64(psb200_ctx *ctx, const uint64_t *x, int64_t n, const uint64_t *keys, int K, int kind,
                                 void *idx, int idx_bytes, psb200_stream stream)
{
    if (!ctx || !x || !keys || !idx || n < 0 || K < 1 || K > 65536 || (kind != 0 && kind != 1) ||
        (idx_bytes != 1 && idx_bytes != 2) || (idx_bytes == 1 && K > 256))
        return fail(PSB200_ERR_INVALID, "index_of64: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n, 256, ctx->sm_count, 16);
    {
        ProfScope ps__(ctx, st, K_SIZEMAP);
        if (idx_bytes == 1 && kind == 0) index_of_kernel<uint8_t, 0><<<g, 256, 0, st>>>(x, n, keys, K, reinterpret_cast<uint8_t *>(idx));
        else if (idx_bytes == 1) index_of_kernel<uint8_t, 1><<<g, 256, 0, st>>>(x, n, keys, K, reinterpret_cast<uint8_t *>(idx));
        else if (kind == 0) index_of_kernel<uint16_t, 0><<<g, 256, 0, st>>>(x, n, keys, K, reinterpret_cast<uint16_t *>(idx));
        else index_of_kernel<uint16_t, 1><<<g, 256, 0, st>>>(x, n, keys, K, reinterpret_cast<uint16_t *>(idx));
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

// ------------------------------------------------------------------------------ drainage
// (drainage_kernels.cuh; reference simulations/_drainage.py:104-154, tools/_sphere_insertions.py:327-385)
static DrainFn make_drain_fn(double c0, double voxel_size, double rho_g, int prec_flags, int64_t inner, bool use_pc)
{
    DrainFn q;
    q.c0 = c0; q.vs64 = voxel_size; q.vs32 = (float)voxel_size; q.rg64 = rho_g; q.rg32 = (float)rho_g;
    q.den64 = (prec_flags & 1) ? 1 : 0; q.h64 = (prec_flags & 2) ? 1 : 0; q.rgh64 = (prec_flags & 4) ? 1 : 0;
    q.use_pc = use_pc ? 1 : 0; q.inner = inner;
    return q;
}

extern "C" int psb200_drain_stats(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user, int64_t n,
                                  int64_t inner, double c0, double voxel_size, double rho_g, int prec_flags,
                                  double *partials, int nblocks, psb200_stream stream)
{
    if (!ctx || !dt || !im || !partials || n < 1 || inner < 1 || nblocks < 1) return fail(PSB200_ERR_INVALID, "drain_stats: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DrainFn q = make_drain_fn(c0, voxel_size, rho_g, prec_flags, inner, pc_user != nullptr);
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_stats_kernel<<<nblocks, 256, 0, st>>>(dt, im, pc_user, n, q, partials);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_drain_threshold(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user,
                                      const uint8_t *residual, int64_t n, int64_t inner, double c0, double voxel_size,
                                      double rho_g, int prec_flags, double p, uint8_t *temp, psb200_stream stream)
{
    if (!ctx || !dt || !im || !temp || n < 1 || inner < 1) return fail(PSB200_ERR_INVALID, "drain_threshold: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DrainFn q = make_drain_fn(c0, voxel_size, rho_g, prec_flags, inner, pc_user != nullptr);
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_threshold_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(dt, im, pc_user, residual, n, q, p, temp);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_drain_newly(psb200_ctx *ctx, const uint8_t *reached, const uint8_t *mask, uint8_t *seeds,
                                  const float *dt, uint16_t *rad, int64_t n, uint64_t *count_dev, int *maxr_dev,
                                  psb200_stream stream)
{
    if (!ctx || !reached || !seeds || !dt || !rad || !count_dev || !maxr_dev || n < 1)
        return fail(PSB200_ERR_INVALID, "drain_newly: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), st));
    CUDA_TRY(cudaMemsetAsync(maxr_dev, 0, sizeof(int), st));
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_newly_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(
            reached, mask, seeds, dt, rad, n, reinterpret_cast<unsigned long long *>(count_dev), maxr_dev);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_drain_classify(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user,
                                     const uint8_t *residual, int64_t n, int64_t inner, double c0, double voxel_size,
                                     double rho_g, int prec_flags, const double *ps_host, int np, uint8_t *cls,
                                     psb200_stream stream)
{
    if (!ctx || !dt || !im || !cls || !ps_host || n < 1 || inner < 1 || np < 1 || np > 254)
        return fail(PSB200_ERR_INVALID, "drain_classify: bad argument");
    for (int i = 1; i < np; ++i)
        if (!(ps_host[i] >= ps_host[i - 1])) return fail(PSB200_ERR_INVALID, "drain_classify: pressures must ascend");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DrainFn q = make_drain_fn(c0, voxel_size, rho_g, prec_flags, inner, pc_user != nullptr);
    DrainPs ps;
    for (int i = 0; i < 254; ++i) ps.p[i] = i < np ? ps_host[i] : 0.0;
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_classify_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(dt, im, pc_user, residual, n, q, ps, np, cls);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_drain_newly_rcls(psb200_ctx *ctx, const uint8_t *rcls, int k, const uint8_t *mask, const float *dt,
                                       uint16_t *rad, int64_t n, uint64_t *count_dev, int *maxr_dev, psb200_stream stream)
{
    if (!ctx || !rcls || !dt || !rad || !count_dev || !maxr_dev || n < 1 || k < 0 || k > 253)
        return fail(PSB200_ERR_INVALID, "drain_newly_rcls: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), st));
    CUDA_TRY(cudaMemsetAsync(maxr_dev, 0, sizeof(int), st));
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_newly_rcls_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(
            rcls, k, mask, dt, rad, n, reinterpret_cast<unsigned long long *>(count_dev), maxr_dev);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" size_t psb200_drain_paint_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx)
{
    if (!ctx) return 0;
    return 2 * align256((size_t)nz * ny * nx * 4) + 512;
}

// Paint the spheres of the newly invaded voxels (rad > 0) into inv with the value `val` (power-diagram passes).
extern "C" int psb200_drain_paint(psb200_ctx *ctx, const uint16_t *rad, int rmax, uint8_t *inv, int val, int64_t nz,
                                  int64_t ny, int64_t nx, void *ws, size_t ws_bytes, psb200_stream stream)
{
    if (!ctx || !rad || !inv || val < 1 || val > 255 || rmax < 0) return fail(PSB200_ERR_INVALID, "drain_paint: bad argument");
    int rc = check_dims("drain_paint", nz, ny, nx);
    if (rc) return rc;
    if (rmax == 0) return PSB200_OK;                         // radius-0 spheres are empty (thresh = r - 0.001 < 0)
    if (rmax > 46340) return fail(PSB200_ERR_UNSUPPORTED, "drain_paint: radius too large");
    const int64_t n = nz * ny * nx;
    char *base = ws ? (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
    if (!base || ws_bytes < 2 * align256((size_t)n * 4) + (size_t)(base - (char *)ws))
        return fail(PSB200_ERR_WORKSPACE, "drain_paint needs %zu workspace bytes", 2 * align256((size_t)n * 4) + 256);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *a = reinterpret_cast<uint32_t *>(base), *b = reinterpret_cast<uint32_t *>(base + align256((size_t)n * 4));
    const uint32_t C = (uint32_t)rmax * (uint32_t)rmax;
    {
        const size_t smem = (size_t)(PX_SEG + 2 * rmax) * 4;
        if ((int)smem > ctx->max_smem_optin) return fail(PSB200_ERR_UNSUPPORTED, "drain_paint: radius too large for the x pass");
        CUDA_TRY(cudaFuncSetAttribute(power_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t jobs = nz * ny * ((nx + PX_SEG - 1) / PX_SEG);
        ProfScope ps__(ctx, st, K_DRAIN);
        power_x_kernel<<<grid_for(jobs, 1, ctx->sm_count, 8), 256, smem, st>>>(rad, a, nz * ny, (int)nx, C, rmax);
    }
    LAUNCH_CHECK(ctx);
    uint32_t *cur = a, *other = b;
    if (ny > 1) {
        rc = launch_minplus<MpSrcU32>(ctx, 1, cur, other, 0, nullptr, nz, ny, nx, st);
        if (rc) return rc;
        uint32_t *t = cur; cur = other; other = t;
    }
    if (nz > 1) {
        rc = launch_minplus<MpSrcU32>(ctx, 0, cur, other, 0, nullptr, nz, ny, nx, st);
        if (rc) return rc;
        uint32_t *t = cur; cur = other; other = t;
    }
    {
        ProfScope ps__(ctx, st, K_DRAIN);
        drain_paint_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, st>>>(cur, inv, n, C, (uint32_t)val);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_set_where_u8(psb200_ctx *ctx, uint8_t *dst, const uint8_t *mask, int value, int64_t n,
                                   psb200_stream stream)
{
    if (!ctx || !dst || !mask || n < 0 || value < 0 || value > 255) return fail(PSB200_ERR_INVALID, "set_where_u8: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        ProfScope ps__(ctx, (cudaStream_t)stream, K_DRAIN);
        set_where_u8_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, (cudaStream_t)stream>>>(dst, mask, (uint8_t)value, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_set_zero_codes_u8(psb200_ctx *ctx, uint8_t *codes, const uint8_t *im, const uint8_t *zero_lut_dev,
                                        int value, int64_t n, psb200_stream stream)
{
    if (!ctx || !codes || !im || !zero_lut_dev || n < 0 || value < 0 || value > 255)
        return fail(PSB200_ERR_INVALID, "set_zero_codes_u8: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        ProfScope ps__(ctx, (cudaStream_t)stream, K_DRAIN);
        set_zero_codes_kernel<<<grid_for(n, 256, ctx->sm_count, 16), 256, 0, (cudaStream_t)stream>>>(codes, im, zero_lut_dev,
                                                                                                 (uint8_t)value, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_mask_pack_u8(psb200_ctx *ctx, const uint8_t *src, uint8_t *bits, int64_t n, psb200_stream stream)
{
    if (!ctx || !src || !bits || n < 0) return fail(PSB200_ERR_INVALID, "mask_pack_u8: bad argument");
    if (n == 0) return PSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb = (n + 7) / 8;
    {
        ProfScope ps__(ctx, st, K_EDT_X);
        mask_pack_kernel<<<grid_for(nb, 256, ctx->sm_count, 16), 256, 0, st>>>(src, bits, nb, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

extern "C" int psb200_mask_unpack_u8(psb200_ctx *ctx, const uint8_t *bits, uint8_t *dst, int64_t n, psb200_stream stream)
{
    if (!ctx || !bits || !dst || n < 0) return fail(PSB200_ERR_INVALID, "mask_unpack_u8: bad argument");
    if (n == 0) return PSB200_OK;
    if (((uintptr_t)dst) & 7u) return fail(PSB200_ERR_INVALID, "mask_unpack_u8: dst must be 8-byte aligned");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb = (n + 7) / 8;
    {
        ProfScope ps__(ctx, st, K_EDT_X);
        mask_unpack_kernel<<<grid_for(nb, 256, ctx->sm_count, 16), 256, 0, st>>>(bits, dst, nb, n);
    }
    LAUNCH_CHECK(ctx);
    return PSB200_OK;
}

