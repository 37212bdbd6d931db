// host_epilogue.cuh -- the last step of local_thickness / porosimetry when the caller wants the
// result in HOST memory (`return imresults`, /root/reference/src/porespy/filters/_funcs.py:1212:
// a float64 array of 8 bytes per voxel).
//
// On the device the result is a one-byte radius INDEX per voxel plus a table of <= 254 radii, so
// shipping float64 over PCIe moves eight times the information.  The epilogue splits the volume:
//   * part A  [0, nA):  the index bytes go over PCIe in chunks; host threads of this library widen
//     every chunk to float64 through the table (a lookup, no arithmetic) with non-temporal stores
//     while the next chunks are still in flight;
//   * part B  [nA, n):  lt_expand_kernel widens on the device into two chunk buffers and the
//     float64 chunks go over PCIe (the path that used to carry everything).
// Both parts run concurrently (PCIe carries nA + 8 (n - nA) bytes, the host memory system absorbs
// 8 nA bytes from its own cores), so the split is a tunable: `cpu_permille` = 1000 nA / n.
// All voxel COMPUTATION (which radius a voxel gets) has happened on the GPU before this point.
#pragma once
#include <emmintrin.h>
#include <thread>
#include <vector>
#include <atomic>
#include "common.cuh"

// out[i] = lut[idx[i]] for i in [0, cnt); `out` 16-byte aligned, cnt even except for a tail.
// (Zeroing the output in the background during the GPU phase and skipping all-zero lines here was
// measured and removed: the epilogue is bound by the host's memory system, and the extra 8.6 GB of
// stores cost more than the 3 GB they saved -- e2e 135 -> 154 ms at 1024^3.)
static void host_widen_slice(const uint8_t *idx, const double *lut, double *out, int64_t cnt)
{
    int64_t i = 0;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        for (; i + 8 <= cnt; i += 8) {
            uint64_t w;
            memcpy(&w, idx + i, 8);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const __m128d v = _mm_set_pd(lut[(w >> (8 * j + 8)) & 0xFFu], lut[(w >> (8 * j)) & 0xFFu]);
                _mm_stream_pd(out + i + j, v);
            }
        }
    }
    for (; i < cnt; ++i) out[i] = lut[idx[i]];
}

struct HostEpilogueStreams {
    cudaStream_t idx_copy = nullptr, f64_copy = nullptr;
};

// Synchronous: returns when out_host[0, n) is complete.
//   stage_host : page-locked, >= nA bytes       ws_dev : device, two float64 chunk buffers
static int host_epilogue_run(psb200_ctx *ctx, HostEpilogueStreams &hs, const uint8_t *idx_dev,
                             const double *lut_host, int nlut, double *out_host, int64_t n,
                             uint8_t *stage_host, size_t stage_bytes, void *ws_dev, size_t ws_bytes,
                             int cpu_permille, int nthreads, cudaStream_t st, int (*launch_expand)(
                                 psb200_ctx *, const uint8_t *, const double *, int, double *, int64_t, cudaStream_t))
{
    if (!hs.idx_copy) CUDA_TRY(cudaStreamCreateWithFlags(&hs.idx_copy, cudaStreamNonBlocking));
    if (!hs.f64_copy) CUDA_TRY(cudaStreamCreateWithFlags(&hs.f64_copy, cudaStreamNonBlocking));
    alignas(64) double lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = i < nlut ? lut_host[i] : 0.0;

    const int64_t CH_A = 1LL << 24;                       // index bytes per part-A chunk (16 MiB -> 128 MiB of float64)
    int64_t nA = (int64_t)((__int128)n * cpu_permille / 1000);
    nA = (nA / CH_A) * CH_A;                              // whole chunks; the rest goes through part B
    if ((size_t)nA > stage_bytes) nA = (int64_t)(stage_bytes / CH_A) * CH_A;
    if (!stage_host || nthreads < 1) nA = 0;
    const int ncA = (int)(nA / CH_A);
    const int64_t nB = n - nA;
    const int64_t CH_B = (int64_t)(ws_bytes / 2 / sizeof(double)) & ~(int64_t)1023;
    if (nB > 0 && (!ws_dev || CH_B < 1024))
        return fail(PSB200_ERR_WORKSPACE, "expand_idx_f64_to_host: device chunk workspace too small (%zu bytes)", ws_bytes);

    cudaEvent_t produced;
    CUDA_TRY(cudaEventCreateWithFlags(&produced, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(produced, st));
    CUDA_TRY(cudaStreamWaitEvent(hs.idx_copy, produced, 0));
    CUDA_TRY(cudaStreamWaitEvent(hs.f64_copy, produced, 0));

    // ---- part A: queue every index chunk, one event per chunk
    std::vector<cudaEvent_t> arrived(ncA);
    for (int c = 0; c < ncA; ++c) {
        CUDA_TRY(cudaEventCreateWithFlags(&arrived[c], cudaEventDisableTiming | cudaEventBlockingSync));
        CUDA_TRY(cudaMemcpyAsync(stage_host + (int64_t)c * CH_A, idx_dev + (int64_t)c * CH_A, (size_t)CH_A,
                                 cudaMemcpyDeviceToHost, hs.idx_copy));
        CUDA_TRY(cudaEventRecord(arrived[c], hs.idx_copy));
    }
    std::atomic<int> thread_err{0};
    std::vector<std::thread> pool;
    const int device = ctx->device;
    if (ncA > 0) {
        pool.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
            pool.emplace_back([=, &arrived, &thread_err, &lut]() {
                if (cudaSetDevice(device) != cudaSuccess) { thread_err = 1; return; }
                const int64_t per = ((CH_A / nthreads) + 7) & ~(int64_t)7;
                const int64_t s = (int64_t)t * per, e = s + per < CH_A ? s + per : CH_A;
                for (int c = 0; c < ncA; ++c) {
                    if (cudaEventSynchronize(arrived[c]) != cudaSuccess) { thread_err = 1; return; }
                    if (s < e)
                        host_widen_slice(stage_host + (int64_t)c * CH_A + s, lut, out_host + (int64_t)c * CH_A + s, e - s);
                }
                _mm_sfence();
            });
        }
    }

    // ---- part B: device-side widening, float64 chunks over PCIe (double-buffered)
    int rc = PSB200_OK;
    cudaEvent_t done[2] = {nullptr, nullptr}, ready[2] = {nullptr, nullptr};
    if (nB > 0) {
        for (int i = 0; i < 2 && rc == PSB200_OK; ++i) {
            if (cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming) != cudaSuccess)
                rc = fail(PSB200_ERR_CUDA, "expand_idx_f64_to_host: event creation failed");
        }
        double *buf[2] = {reinterpret_cast<double *>(ws_dev), reinterpret_cast<double *>(ws_dev) + CH_B};
        int i = 0;
        for (int64_t s = nA; s < n && rc == PSB200_OK; s += CH_B, ++i) {
            const int64_t cnt = s + CH_B < n ? CH_B : n - s;
            const int b = i & 1;
            if (i >= 2 && cudaStreamWaitEvent(st, done[b], 0) != cudaSuccess) { rc = fail(PSB200_ERR_CUDA, "wait"); break; }
            rc = launch_expand(ctx, idx_dev + s, lut_host, nlut, buf[b], cnt, st);
            if (rc) break;
            if (cudaEventRecord(ready[b], st) != cudaSuccess || cudaStreamWaitEvent(hs.f64_copy, ready[b], 0) != cudaSuccess ||
                cudaMemcpyAsync(out_host + s, buf[b], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, hs.f64_copy) != cudaSuccess ||
                cudaEventRecord(done[b], hs.f64_copy) != cudaSuccess)
                rc = fail(PSB200_ERR_CUDA, "expand_idx_f64_to_host: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (auto &th : pool) th.join();
    cudaError_t e1 = cudaStreamSynchronize(hs.idx_copy), e2 = cudaStreamSynchronize(hs.f64_copy);
    cudaError_t e3 = cudaStreamSynchronize(st);
    for (auto &ev : arrived) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (done[i]) cudaEventDestroy(done[i]);
        if (ready[i]) cudaEventDestroy(ready[i]);
    }
    cudaEventDestroy(produced);
    if (rc) return rc;
    if (thread_err) return fail(PSB200_ERR_CUDA, "expand_idx_f64_to_host: a host thread could not wait for its chunk");
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(PSB200_ERR_CUDA, "expand_idx_f64_to_host: %s",
                    cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return PSB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Host prologue: the input volume of the reference API is one byte per voxel (numpy bool / uint8),
// of which the kernels use one bit (`byte != 0`, F:1126 `im > 0`).  Host threads pack the volume to
// bits chunk by chunk (a compare and a movemask per 16 voxels), every finished chunk crosses PCIe
// at an eighth of the size while the next ones are being packed, and one kernel on the device
// spreads the bits back to 0/1 bytes.
__global__ void __launch_bounds__(256)
mask_unpack_kernel(const uint8_t *__restrict__ bits, uint8_t *__restrict__ out, int64_t nbytes_bits, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes_bits; i += step) {
        const uint32_t b = bits[i];
        // bit j of b -> byte j (0 / 1)
        const uint32_t lo = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
        const uint32_t hi = ((b >> 4) * 0x00204081u) & 0x01010101u;
        if (8 * i + 8 <= n) *reinterpret_cast<uint2 *>(out + 8 * i) = make_uint2(lo, hi);
        else
            for (int j = 0; j < 8 && 8 * i + j < n; ++j) out[8 * i + j] = (uint8_t)((b >> j) & 1u);
    }
}

// the device-side counterpart (z-slab shards send their EDT halo planes as bits): bits[i] bit j = (src[8 i + j] != 0)
__global__ void __launch_bounds__(256)
mask_pack_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ bits, int64_t nbytes_bits, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = (((uintptr_t)src) & 7u) == 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes_bits; i += step) {
        uint32_t b = 0;
        if (aligned && 8 * i + 8 <= n) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src + 8 * i));
            // non-zero byte -> its bit 7 set (carry-free SWAR), then gather the four bit-7s of a word into a nibble
            const uint32_t lo = ((v.x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v.x, hi = ((v.y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v.y;
            b = ((((lo >> 7) & 0x01010101u) * 0x01020408u) >> 24) | (((((hi >> 7) & 0x01010101u) * 0x01020408u) >> 24) << 4);
        } else {
            for (int j = 0; j < 8 && 8 * i + j < n; ++j) b |= (src[8 * i + j] != 0 ? 1u : 0u) << j;
        }
        bits[i] = (uint8_t)b;
    }
}

// bits[i] bit j = (src[8 i + j] != 0) for the voxels [v0, v1); v0 % 8 == 0
static void host_pack_slice(const uint8_t *src, uint8_t *bits, int64_t v0, int64_t v1)
{
    int64_t v = v0;
    const __m128i zero = _mm_setzero_si128();
    for (; v + 16 <= v1; v += 16) {
        const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + v));
        const uint32_t m = ~(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(x, zero)) & 0xFFFFu;
        bits[v >> 3] = (uint8_t)(m & 0xFFu);
        bits[(v >> 3) + 1] = (uint8_t)(m >> 8);
    }
    for (; v < v1; v += 8) {
        uint32_t m = 0;
        for (int j = 0; j < 8 && v + j < v1; ++j) m |= (src[v + j] != 0 ? 1u : 0u) << j;
        bits[v >> 3] = (uint8_t)m;
    }
}

// Synchronous with respect to the host source AND the staging buffer (both may be reused on return);
// the device side (unpack kernel) is ordered on `st`.  stage_host: page-locked, >= ceil(n / 8) bytes; bits_dev: device, same size.
static int host_upload_mask(psb200_ctx *ctx, const uint8_t *src_host, int64_t n, uint8_t *dst_dev,
                            uint8_t *stage_host, uint8_t *bits_dev, int nthreads, cudaStream_t st)
{
    const int64_t CH = 1LL << 23;                          // voxels per chunk (8 MiB in, 1 MiB of bits)
    const int nch = (int)((n + CH - 1) / CH);
    std::vector<std::atomic<int>> done(nch);
    for (auto &d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&]() {
            for (int c = next.fetch_add(1); c < nch; c = next.fetch_add(1)) {
                const int64_t v0 = (int64_t)c * CH, v1 = v0 + CH < n ? v0 + CH : n;
                host_pack_slice(src_host, stage_host, v0, v1);
                done[c].store(1, std::memory_order_release);
            }
        });
    cudaError_t err = cudaSuccess;
    for (int c = 0; c < nch; ++c) {
        while (!done[c].load(std::memory_order_acquire)) std::this_thread::yield();
        const int64_t v0 = (int64_t)c * CH, v1 = v0 + CH < n ? v0 + CH : n;
        const int64_t b0 = v0 >> 3, b1 = (v1 + 7) >> 3;
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(bits_dev + b0, stage_host + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st);
    }
    for (auto &th : pool) th.join();
    if (err != cudaSuccess) return fail(PSB200_ERR_CUDA, "upload_mask_u8: %s", cudaGetErrorString(err));
    // the copies read the staging buffer asynchronously: wait until the last one has left host memory, so
    // that the caller may repack the same buffer for its next volume right away (fg, inlets, outlets are
    // uploaded back to back by the flood users)
    cudaEvent_t copied;
    if ((err = cudaEventCreateWithFlags(&copied, cudaEventDisableTiming)) == cudaSuccess) {
        err = cudaEventRecord(copied, st);
        if (err == cudaSuccess) err = cudaEventSynchronize(copied);
        cudaEventDestroy(copied);
    }
    if (err != cudaSuccess) return fail(PSB200_ERR_CUDA, "upload_mask_u8: %s", cudaGetErrorString(err));
    const int64_t nb = (n + 7) >> 3;
    int g = (int)((nb + 255) / 256);
    if (g > ctx->sm_count * 16) g = ctx->sm_count * 16;
    mask_unpack_kernel<<<g, 256, 0, st>>>(bits_dev, dst_dev, nb, n);
    ctx->launches++;
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(PSB200_ERR_CUDA, "upload_mask_u8: %s", cudaGetErrorString(err));
    return PSB200_OK;
}
