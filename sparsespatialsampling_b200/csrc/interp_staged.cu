// Staged variant of the export interpolation (fp32 fast path), see interp.cu for the operator.
//
// The unique source rows of a tile of 32 cells are copied ONCE into shared memory with 1-D TMA bulk copies
// (cp.async.bulk, one per row segment, all completing on one mbarrier), then every cell of the tile reads its
// k neighbours from shared memory.  ncu on the direct kernel showed why this matters: neighbouring cells share
// most of their source rows (C2: 5.4 references per unique row), but with ~1800 CTAs each walking 32 cells the
// re-use distance exceeded L2, so DRAM read 2.1x the unique bytes and L2->SM carried all k references.  Here a
// tile's rows cross L2->SM once, the CTAs in flight cover a compact Morton window (tens of MB), and
// neighbouring tiles find their shared rows in L2.
#include "common.cuh"
#include "tma.cuh"
#include "../../include/s3b200.h"

namespace s3 {

constexpr int kTileCells = 32;
constexpr int kTileThreads = 256;
constexpr int kTileMaxRefs = 2048;  // kTileCells * k rounded up to a power of two must fit
int g_staging = 0;                  // 0 = TMA bulk copies, 1 = cp.async (s3_set_tuning key 1)
int g_stage_budget_kb = 56;         // shared memory per CTA for staged rows (s3_set_tuning key 2)

// Tile preprocessing (once per KNN cache): unique source rows of the tile + local index of every reference.
__global__ void __launch_bounds__(kTileThreads)
tile_build_kernel(const int32_t* __restrict__ idx, int64_t n_cells, int k, int refs_pow2,
                  int32_t* __restrict__ tile_rows, int32_t* __restrict__ tile_nrows, uint16_t* __restrict__ tile_lidx) {
    extern __shared__ uint32_t s_keys[];          // refs_pow2 sorted keys, then refs_pow2 unique keys
    uint32_t* s_uniq = s_keys + refs_pow2;
    __shared__ int s_chunk[kTileThreads + 1];
    const int64_t tile = blockIdx.x;
    const int64_t cell0 = tile * kTileCells;
    const int ncell = (int)((n_cells - cell0) < kTileCells ? (n_cells - cell0) : kTileCells);
    const int nref = ncell * k;
    const int cap = kTileCells * k;
    for (int i = threadIdx.x; i < refs_pow2; i += kTileThreads)
        s_keys[i] = i < nref ? (uint32_t)idx[cell0 * k + i] : 0xffffffffu;
    __syncthreads();
    // bitonic sort, ascending
    for (int size = 2; size <= refs_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < refs_pow2 / 2; i += kTileThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint32_t a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == up) { s_keys[lo] = b; s_keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    // compact the run heads: per-thread chunk counts -> prefix -> write
    const int per = (refs_pow2 + kTileThreads - 1) / kTileThreads;
    const int lo = threadIdx.x * per, hi = min(lo + per, refs_pow2);
    int heads = 0;
    for (int i = lo; i < hi; ++i)
        heads += (s_keys[i] != 0xffffffffu && (i == 0 || s_keys[i] != s_keys[i - 1])) ? 1 : 0;
    s_chunk[threadIdx.x + 1] = heads;
    if (threadIdx.x == 0) s_chunk[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i <= kTileThreads; ++i) s_chunk[i] += s_chunk[i - 1];
    __syncthreads();
    int pos = s_chunk[threadIdx.x];
    for (int i = lo; i < hi; ++i)
        if (s_keys[i] != 0xffffffffu && (i == 0 || s_keys[i] != s_keys[i - 1])) s_uniq[pos++] = s_keys[i];
    __syncthreads();
    const int nuniq = s_chunk[kTileThreads];
    if (threadIdx.x == 0) tile_nrows[tile] = nuniq;
    for (int i = threadIdx.x; i < cap; i += kTileThreads) tile_rows[tile * cap + i] = i < nuniq ? (int32_t)s_uniq[i] : 0;
    // local index of every reference: binary search in the unique list
    for (int i = threadIdx.x; i < cap; i += kTileThreads) {
        uint16_t li = 0;
        if (i < nref) {
            const uint32_t key = (uint32_t)idx[cell0 * k + i];
            int a = 0, b = nuniq - 1;
            while (a < b) {
                const int m = (a + b) >> 1;
                if (s_uniq[m] < key) a = m + 1; else b = m;
            }
            li = (uint16_t)a;
        }
        tile_lidx[tile * cap + i] = li;
    }
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int W, int STAGING>   // W = staged columns per row (floats), multiple of 128; STAGING 0 = TMA bulk, 1 = cp.async
__global__ void __launch_bounds__(kTileThreads)
interp_staged_kernel(const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ tile_rows,
                     const int32_t* __restrict__ tile_nrows, const uint16_t* __restrict__ tile_lidx,
                     const float* __restrict__ w, int64_t n_cells, int k, int n_chunks, int r_smem,
                     const int32_t* __restrict__ out_row, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int cap = kTileCells * k;
    float* s_rows = reinterpret_cast<float*>(smem_raw);                                   // [r_smem][W]
    float* s_w = reinterpret_cast<float*>(smem_raw + (size_t)r_smem * W * sizeof(float));  // [cap]
    uint16_t* s_lidx = reinterpret_cast<uint16_t*>(s_w + cap);                             // [cap]

    const int64_t tile = blockIdx.x / n_chunks;
    const int chunk = (int)(blockIdx.x % n_chunks);
    const int64_t cell0 = tile * kTileCells;
    const int ncell = (int)((n_cells - cell0) < kTileCells ? (n_cells - cell0) : kTileCells);
    const int64_t col0 = (int64_t)chunk * W;
    const int wcur = (int)((row_len - col0) < W ? (row_len - col0) : W);
    const int nrows = tile_nrows[tile];
    const int nstage = nrows < r_smem ? nrows : r_smem;
    const uint32_t row_bytes = (uint32_t)wcur * 4u;
    const int32_t* rows = tile_rows + tile * cap;

    if (STAGING == 0) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, (uint32_t)nstage * row_bytes + (uint32_t)cap * 6u);
            tma_load_1d(s_w, w + tile * cap, (uint32_t)cap * 4u, &bar);
            tma_load_1d(s_lidx, tile_lidx + tile * cap, (uint32_t)cap * 2u, &bar);
        }
        __syncthreads();
        for (int r = threadIdx.x; r < nstage; r += kTileThreads)
            tma_load_1d(s_rows + (size_t)r * W, data + (int64_t)rows[r] * row_len + col0, row_bytes, &bar);
        mbar_wait(&bar, 0);
    } else {
        // 16-byte cp.async pieces issued by all threads; a warp covers 512 contiguous bytes of one row
        const int ppr = wcur >> 2;                       // 16-byte pieces per row
        const int total = nstage * ppr;
        for (int p = threadIdx.x; p < total; p += kTileThreads) {
            const int r = p / ppr, c4 = p - r * ppr;
            cp_async_16(s_rows + (size_t)r * W + c4 * 4, data + (int64_t)rows[r] * row_len + col0 + c4 * 4);
        }
        for (int i = threadIdx.x; i < cap; i += kTileThreads) {
            s_w[i] = w[tile * cap + i];
            s_lidx[i] = tile_lidx[tile * cap + i];
        }
        cp_async_commit_wait_all();
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int S = W / 128;
    for (int c = warp; c < ncell; c += kTileThreads / 32) {
        float4 acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint16_t* li = s_lidx + c * k;
        const float* cw = s_w + c * k;
        for (int j = 0; j < k; ++j) {
            const int r = li[j];
            const float wj = cw[j];
            if (r < nstage) {
                const float4* src = reinterpret_cast<const float4*>(s_rows + (size_t)r * W) + lane;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const float4 x = src[s * 32];
                    acc[s].x = fmaf(wj, x.x, acc[s].x); acc[s].y = fmaf(wj, x.y, acc[s].y);
                    acc[s].z = fmaf(wj, x.z, acc[s].z); acc[s].w = fmaf(wj, x.w, acc[s].w);
                }
            } else {
                // tile has more unique rows than the staging buffer holds: read this one directly
                const float* g = data + (int64_t)rows[r] * row_len + col0;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int cc = s * 128 + lane * 4;
                    if (cc < wcur) {
                        const float4 x = *reinterpret_cast<const float4*>(g + cc);
                        acc[s].x = fmaf(wj, x.x, acc[s].x); acc[s].y = fmaf(wj, x.y, acc[s].y);
                        acc[s].z = fmaf(wj, x.z, acc[s].z); acc[s].w = fmaf(wj, x.w, acc[s].w);
                    }
                }
            }
        }
        const int64_t orow = out_row ? (int64_t)out_row[cell0 + c] : (cell0 + c);
        float* o = out + orow * row_len + col0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int cc = s * 128 + lane * 4;
            if (cc < wcur) *reinterpret_cast<float4*>(o + cc) = acc[s];
        }
    }
}

}  // namespace s3

using namespace s3;

extern "C" int s3_interp_tiles_build(const int32_t* d_idx, int64_t n_cells, int k, int32_t* d_tile_rows,
                                     int32_t* d_tile_nrows, uint16_t* d_tile_lidx, void* stream) {
    S3_REQUIRE(d_idx && d_tile_rows && d_tile_nrows && d_tile_lidx, "s3_interp_tiles_build: NULL argument");
    S3_REQUIRE(k >= 1 && kTileCells * k <= kTileMaxRefs, "s3_interp_tiles_build: k=%d too large", k);
    if (n_cells == 0) return S3_OK;
    int p2 = 1;
    while (p2 < kTileCells * k) p2 <<= 1;
    const int64_t tiles = ceil_div(n_cells, kTileCells);
    tile_build_kernel<<<(unsigned)tiles, kTileThreads, 2 * p2 * sizeof(uint32_t), (cudaStream_t)stream>>>(
        d_idx, n_cells, k, p2, d_tile_rows, d_tile_nrows, d_tile_lidx);
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}

extern "C" int s3_interp_staged(const float* d_data, int64_t n_src, int64_t row_len, const int32_t* d_tile_rows,
                                const int32_t* d_tile_nrows, const uint16_t* d_tile_lidx, const float* d_w,
                                int64_t n_cells, int k, int max_rows, int chunk_cols, const int32_t* d_out_row,
                                float* d_out, void* stream) {
    S3_REQUIRE(d_data && d_tile_rows && d_tile_nrows && d_tile_lidx && d_w && d_out, "s3_interp_staged: NULL argument");
    S3_REQUIRE(row_len % 4 == 0 && ((uintptr_t)d_data % 16) == 0 && ((uintptr_t)d_out % 16) == 0,
               "s3_interp_staged: rows must be 16-byte aligned (row_len %% 4 == 0)");
    S3_REQUIRE(chunk_cols == 128 || chunk_cols == 256, "s3_interp_staged: chunk_cols must be 128 or 256");
    S3_REQUIRE(kTileCells * k <= kTileMaxRefs, "s3_interp_staged: k too large");
    (void)n_src;
    if (n_cells == 0 || row_len == 0) return S3_OK;
    const int cap = kTileCells * k;
    const size_t table_bytes = (size_t)cap * 6;
    const size_t budget = (size_t)g_stage_budget_kb * 1024;
    int r_smem = max_rows < 1 ? 1 : max_rows;
    const size_t max_fit = (budget - table_bytes) / ((size_t)chunk_cols * 4);
    if ((size_t)r_smem > max_fit) r_smem = (int)max_fit;
    const size_t smem = (size_t)r_smem * chunk_cols * 4 + table_bytes;
    const int n_chunks = (int)ceil_div(row_len, chunk_cols);
    const int64_t tiles = ceil_div(n_cells, kTileCells);
    const int64_t blocks = tiles * n_chunks;
    S3_REQUIRE(blocks < ((int64_t)1 << 31), "s3_interp_staged: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
#define S3_LAUNCH_STAGED(WW, SS)                                                                                  \
    do {                                                                                                         \
        S3_CUDA(cudaFuncSetAttribute(interp_staged_kernel<WW, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                     (int)smem));                                                                \
        interp_staged_kernel<WW, SS><<<(unsigned)blocks, kTileThreads, smem, st>>>(                               \
            d_data, row_len, d_tile_rows, d_tile_nrows, d_tile_lidx, d_w, n_cells, k, n_chunks, r_smem, d_out_row, \
            d_out);                                                                                              \
    } while (0)
    if (chunk_cols == 128 && g_staging == 0) S3_LAUNCH_STAGED(128, 0);
    else if (chunk_cols == 128) S3_LAUNCH_STAGED(128, 1);
    else if (g_staging == 0) S3_LAUNCH_STAGED(256, 0);
    else S3_LAUNCH_STAGED(256, 1);
#undef S3_LAUNCH_STAGED
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}
