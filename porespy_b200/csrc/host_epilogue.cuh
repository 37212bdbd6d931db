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
// prezeroed: out already holds 0.0 everywhere (psb200_host_zero_begin), so a group of 8 voxels whose
// indices are all 0 (solid, or never invaded: lut[0] == 0.0) needs no store -- on a porous volume
// about a third of the 64-byte lines.
static void host_widen_slice(const uint8_t *idx, const double *lut, double *out, int64_t cnt, bool prezeroed)
{
    int64_t i = 0;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        for (; i + 8 <= cnt; i += 8) {
            uint64_t w;
            memcpy(&w, idx + i, 8);
            if (prezeroed && w == 0) continue;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const __m128d v = _mm_set_pd(lut[(w >> (8 * j + 8)) & 0xFFu], lut[(w >> (8 * j)) & 0xFFu]);
                _mm_stream_pd(out + i + j, v);
            }
        }
    }
    for (; i < cnt; ++i) out[i] = lut[idx[i]];
}

// ---- output buffer zeroed in the background while the GPU computes
struct HostZeroJob {
    std::vector<std::thread> pool;
};

static void host_zero_slice(double *out, int64_t cnt)
{
    int64_t i = 0;
    while (i < cnt && (reinterpret_cast<uintptr_t>(out + i) & 15u)) out[i++] = 0.0;
    const __m128d z = _mm_setzero_pd();
    for (; i + 8 <= cnt; i += 8) {
        _mm_stream_pd(out + i, z); _mm_stream_pd(out + i + 2, z);
        _mm_stream_pd(out + i + 4, z); _mm_stream_pd(out + i + 6, z);
    }
    for (; i < cnt; ++i) out[i] = 0.0;
    _mm_sfence();
}

struct HostEpilogueStreams {
    cudaStream_t idx_copy = nullptr, f64_copy = nullptr;
};

// Synchronous: returns when out_host[0, n) is complete.
//   stage_host : page-locked, >= nA bytes       ws_dev : device, two float64 chunk buffers
static int host_epilogue_run(psb200_ctx *ctx, HostEpilogueStreams &hs, const uint8_t *idx_dev,
                             const double *lut_host, int nlut, double *out_host, int64_t n,
                             uint8_t *stage_host, size_t stage_bytes, void *ws_dev, size_t ws_bytes,
                             int cpu_permille, int nthreads, bool prezeroed, cudaStream_t st, int (*launch_expand)(
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
                const bool pz = prezeroed && lut[0] == 0.0;
                if (cudaSetDevice(device) != cudaSuccess) { thread_err = 1; return; }
                const int64_t per = ((CH_A / nthreads) + 7) & ~(int64_t)7;
                const int64_t s = (int64_t)t * per, e = s + per < CH_A ? s + per : CH_A;
                for (int c = 0; c < ncA; ++c) {
                    if (cudaEventSynchronize(arrived[c]) != cudaSuccess) { thread_err = 1; return; }
                    if (s < e)
                        host_widen_slice(stage_host + (int64_t)c * CH_A + s, lut, out_host + (int64_t)c * CH_A + s, e - s, pz);
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
