// Hand-written stable LSD radix sort for (u64 key, u32 value) pairs.
//
// Used by (i) the Morton ordering of the original point cloud in the KNN index build,
// (ii) the ordering of the cells selected for refinement (gain descending, index ascending;
// reference: sparseSpatialSampling/s_cube.py:601-602) and (iii) the node de-duplication.
//
// 8-bit digits, three kernels per pass (tile histogram -> exclusive scan -> stable scatter).
// Stability inside a tile comes from the warp-contiguous element order (warp, item, lane) and
// __match_any_sync ranking, so equal keys keep their input order.
#pragma once
#include "common.cuh"

namespace s3 {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kSortWarps = kSortThreads / 32;

static __global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist,
                  int nblocks) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = base + (int64_t)warp * (32 * kSortItems);
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        int64_t i = wbase + j * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(keys[i] >> shift) & 255u;
            atomicAdd(&sh[d], 1u);
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan of `count` u32 entries, single block (count = 256 * nblocks is small).
static __global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ data, int64_t count) {
    __shared__ uint32_t partial[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (count + 1023) / 1024;
    const int64_t lo = (int64_t)t * chunk;
    const int64_t hi = lo + chunk < count ? lo + chunk : count;
    uint32_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += data[i];
    partial[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = (t >= off) ? partial[t - off] : 0u;
        __syncthreads();
        partial[t] += v;
        __syncthreads();
    }
    uint32_t run = (t == 0) ? 0u : partial[t - 1];
    for (int64_t i = lo; i < hi; ++i) {
        uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
}

static __global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     const uint32_t* __restrict__ offsets, int nblocks) {
    __shared__ uint32_t whist[kSortWarps][257];
    for (int i = threadIdx.x; i < kSortWarps * 257; i += kSortThreads) (&whist[0][0])[i] = 0;
    __syncthreads();

    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = base + (int64_t)warp * (32 * kSortItems);
    const uint32_t lt_mask = (1u << lane) - 1u;

    uint64_t k[kSortItems];
    uint32_t v[kSortItems];
    uint32_t dg[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        int64_t i = wbase + j * 32 + lane;
        if (i < n) {
            k[j] = keys_in[i];
            v[j] = vals_in[i];
            dg[j] = (uint32_t)(k[j] >> shift) & 255u;
        } else {
            k[j] = 0;
            v[j] = 0;
            dg[j] = 256u;
        }
    }
    // phase 1: per-warp digit counts
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        uint32_t m = __match_any_sync(0xffffffffu, dg[j]);
        if ((m & lt_mask) == 0) whist[warp][dg[j]] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: per digit, turn the warp counts into starting offsets (global + earlier warps)
    {
        const int d = threadIdx.x;  // 256 threads == 256 digits
        uint32_t run = offsets[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            uint32_t c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: stable scatter
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        uint32_t m = __match_any_sync(0xffffffffu, dg[j]);
        uint32_t rank = __popc(m & lt_mask);
        uint32_t start = whist[warp][dg[j]];
        __syncwarp();
        if (rank == 0) whist[warp][dg[j]] = start + __popc(m);
        __syncwarp();
        if (dg[j] < 256u) {
            uint32_t pos = start + rank;
            keys_out[pos] = k[j];
            vals_out[pos] = v[j];
        }
    }
}

// Sorts ascending by the key bits [begin_bit, end_bit). Ping-pongs between (a) and (b);
// *in_a tells where the result ended up. n must be < 2^31.
static inline int radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                                   int64_t n, int begin_bit, int end_bit, cudaStream_t stream, bool* in_a) {
    *in_a = true;
    if (n <= 1) return S3_OK;
    S3_REQUIRE(n < (int64_t)1 << 31, "radix_sort_pairs: n=%lld too large", (long long)n);
    const int nblocks = (int)ceil_div(n, kSortTile);
    Scratch scratch(stream);
    uint32_t* hist = nullptr;
    S3_TRY(scratch.alloc(&hist, (size_t)256 * nblocks));
    uint64_t* kin = keys_a;
    uint32_t* vin = vals_a;
    uint64_t* kout = keys_b;
    uint32_t* vout = vals_b;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        radix_hist_kernel<<<nblocks, kSortThreads, 0, stream>>>(kin, n, shift, hist, nblocks);
        radix_scan_kernel<<<1, 1024, 0, stream>>>(hist, (int64_t)256 * nblocks);
        radix_scatter_kernel<<<nblocks, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, hist, nblocks);
        S3_LAUNCH_CHECK();
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
        *in_a = !*in_a;
    }
    return S3_OK;
}

}  // namespace s3
