// sizemap_kernels.cuh -- post-processing of the radius map ON ITS INDEX FORM:
//   size_to_seq / size_to_satn / seq_to_satn   /root/reference/src/porespy/filters/_size_seq_satn.py:16-221
//   pore_size_distribution                     /root/reference/src/porespy/metrics/_funcs.py:558-632
//   pc_curve (sizes branch)                    /root/reference/src/porespy/metrics/_funcs.py:1073-1090
//
// The map porosimetry / local_thickness return holds at most 254 distinct values (one per radius), and
// the device already has it as one index byte per voxel plus a table.  Every function above is a
// composition of (a) set operations on the distinct values (unique, digitize, rank, make_contiguous),
// (b) counts of voxels per value (optionally split by a pore mask `im`), and (c) a pointwise map
// value -> new value.  So the device work is one histogram of the index map and one table expansion; the
// arithmetic on the <= 2 * 65536 (value, mask) combinations is the reference's own numpy code on the host.
// Arbitrary host arrays enter through a distinct-value hash + index-of pass (idx_build kernels).
#pragma once
#include "common.cuh"

// counts[(mask && mask[i]) ? K + idx[i] : idx[i]] += 1      (counts: 2K or K unsigned 64-bit, zeroed by the caller)
#define HIST_SMEM_BINS 8192
template <typename I>
__global__ void __launch_bounds__(256)
hist_idx_kernel(const I *__restrict__ idx, const uint8_t *__restrict__ mask, int64_t n, int K,
                unsigned long long *__restrict__ counts)
{
    __shared__ uint32_t sh[HIST_SMEM_BINS];
    const int nb = mask ? 2 * K : K;
    const bool use_sh = nb <= HIST_SMEM_BINS;
    if (use_sh)
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        int b = (int)idx[i];
        if (mask && mask[i]) b += K;
        if (use_sh) atomicAdd(&sh[b], 1u);              // < 2^32 voxels per block
        else atomicAdd(&counts[b], 1ull);
    }
    __syncthreads();
    if (use_sh)
        for (int i = threadIdx.x; i < nb; i += blockDim.x)
            if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
}

// out[i] = lut[(mask && mask[i]) ? K + idx[i] : idx[i]]     8-byte payload (float64 or int64 bit patterns)
template <typename I>
__global__ void __launch_bounds__(256)
expand_lut8_kernel(const I *__restrict__ idx, const uint8_t *__restrict__ mask, const uint64_t *__restrict__ lut,
                   uint64_t *__restrict__ out, int64_t n, int K)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        int b = (int)idx[i];
        if (mask && mask[i]) b += K;
        out[i] = __ldg(lut + b);
    }
}

// out[i] = lut[(mask && mask[i]) ? K + idx[i] : idx[i]]     one-byte payload (masks such as `seq >= i`)
template <typename I>
__global__ void __launch_bounds__(256)
expand_lut1_kernel(const I *__restrict__ idx, const uint8_t *__restrict__ mask, const uint8_t *__restrict__ lut,
                   uint8_t *__restrict__ out, int64_t n, int K)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        int b = (int)idx[i];
        if (mask && mask[i]) b += K;
        out[i] = __ldg(lut + b);
    }
}

// ---- arbitrary 8-byte arrays -> index form
// Distinct bit patterns of x[] into an open-addressing table (cap = power of two).  The empty marker is
// 0x8000...0: as float64 it is -0.0, which the caller canonicalises to +0.0 beforehand (numpy's `unique` treats
// them as one value anyway); as int64 it is INT64_MIN, which no sequence / label map holds.
#define DISTINCT_EMPTY 0x8000000000000000ull
__global__ void __launch_bounds__(256)
fill_u64_kernel(unsigned long long *__restrict__ p, uint32_t n, unsigned long long v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
__device__ __forceinline__ uint32_t hash64(uint64_t k)
{
    k ^= k >> 33; k *= 0xFF51AFD7ED558CCDull; k ^= k >> 33; k *= 0xC4CEB9FE1A85EC53ull; k ^= k >> 33;
    return (uint32_t)k;
}

__global__ void __launch_bounds__(256)
distinct64_kernel(const uint64_t *__restrict__ x, int64_t n, unsigned long long *__restrict__ table, uint32_t cap,
                  int *__restrict__ overflow)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    uint64_t last = DISTINCT_EMPTY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const uint64_t key = x[i];
        if (key == last) continue;                      // runs of equal values are the common case
        last = key;
        uint32_t h = hash64(key) & (cap - 1);
        uint32_t probes = 0;
        while (true) {
            const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&table[h]);
            if (cur == key) break;
            if (cur == DISTINCT_EMPTY) {
                const unsigned long long old = atomicCAS(&table[h], DISTINCT_EMPTY, (unsigned long long)key);
                if (old == DISTINCT_EMPTY || old == key) break;
            }
            h = (h + 1) & (cap - 1);
            if (++probes >= cap) { *overflow = 1; break; }
        }
    }
}

// idx[i] = position of x[i] in the sorted table keys[0..K) (every x[i] is present).  KIND 0: keys compare as
// float64, 1: as int64.
template <typename I, int KIND>
__global__ void __launch_bounds__(256)
index_of_kernel(const uint64_t *__restrict__ x, int64_t n, const uint64_t *__restrict__ keys, int K, I *__restrict__ idx)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const uint64_t key = x[i];
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const uint64_t kv = __ldg(keys + mid);
            bool less;                                   // keys[mid] < key
            if (KIND == 0) less = __longlong_as_double((long long)kv) < __longlong_as_double((long long)key);
            else less = (long long)kv < (long long)key;
            if (less) lo = mid + 1; else hi = mid;
        }
        idx[i] = (I)lo;
    }
}
