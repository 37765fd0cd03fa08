// Pipelined, persistent variant of the staged export interpolation (fp32 fast path).
//
// Same tile structures as interp_staged.cu (unique source rows of every 32-cell tile + local indices), but the
// load -> wait -> compute chain is broken up: each CTA is persistent (one per SM), owns a ring of kStages shared
// memory stages and is warp-specialised --
//   * warp 8 (producer): for every work item (tile, column chunk) of this CTA it waits until the stage is free,
//     arms the stage's mbarrier with the byte count and issues one 1-D TMA bulk copy per unique row segment
//     (cp.async.bulk.shared::cluster.global, 32 lanes issue in parallel) plus two for the (lidx, w) tables;
//   * warps 0..7 (consumers): wait for the stage, every warp interpolates 4 of the tile's 32 cells reading the k
//     neighbour segments from shared memory (LDS.128, conflict free), stores 128-bit results, releases the stage.
// Rows therefore cross L2->SM once per tile instead of once per reference and never pass the LSU on the way in,
// while the TMA engine keeps fetching the next tiles during the compute.
#include <cuda.h>
#include "common.cuh"
#include "tma.cuh"
#include "../../include/s3b200.h"

namespace s3 {

constexpr int kPipeTileCells = 32;
constexpr int kPipeConsumerWarps = 8;
constexpr int kPipeThreads = (kPipeConsumerWarps + 1) * 32;
constexpr int kPipeMaxStages = 8;
int g_pipe_prefetch = 0;            // work items of L2 prefetch distance (s3_set_tuning key 6)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Blackwell TMA gather: four rows of a 2-D tensor (row indices r0..r3, column start c0, box = W x 1) in ONE request;
// lands as 4 consecutive [W]-float rows in shared memory. Cuts the TMA request count 4x against 1-D bulk copies,
// which measured ~100-130 cycles per request and SM regardless of size.
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

template <int W, bool G4>
__global__ void __launch_bounds__(kPipeThreads, 1)
interp_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ tile_rows,
                   const int32_t* __restrict__ tile_nrows, const uint16_t* __restrict__ tile_lidx,
                   const float* __restrict__ w, int64_t n_cells, int k, int n_chunks, int64_t n_items, int r_smem,
                   int n_stages, int prefetch_items, const int32_t* __restrict__ out_row, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kPipeMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kPipeMaxStages];
    const int cap = kPipeTileCells * k;
    const size_t stage_bytes = (((size_t)r_smem * W * 4 + (size_t)cap * 6) + 127) & ~(size_t)127;   // rows | w | lidx
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kPipeConsumerWarps);
        }
    }
    __syncthreads();

    if (warp == kPipeConsumerWarps) {
        // ------------------------------------------------------------------ producer warp
        int64_t it = 0;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int s = (int)(it % n_stages);
            const uint32_t round = (uint32_t)(it / n_stages);
            if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1u);
            const int64_t tile = item / n_chunks;
            const int chunk = (int)(item % n_chunks);
            const int64_t col0 = (int64_t)chunk * W;
            const int wcur = (int)((row_len - col0) < W ? (row_len - col0) : W);
            const int nrows = tile_nrows[tile];
            const int nstage = nrows < r_smem ? nrows : r_smem;
            const uint32_t row_bytes = (uint32_t)wcur * 4u;
            unsigned char* base = smem_raw + (size_t)s * stage_bytes;
            float* s_rows = reinterpret_cast<float*>(base);
            float* s_w = reinterpret_cast<float*>(base + (size_t)r_smem * W * 4);
            uint16_t* s_lidx = reinterpret_cast<uint16_t*>(s_w + cap);
            const int32_t* rows = tile_rows + tile * cap;
            if (prefetch_items > 0) {
                // pull the rows of a later work item into L2 now, so that its TMA loads find them there
                const int64_t pitem = item + (int64_t)prefetch_items * gridDim.x;
                if (pitem < n_items) {
                    const int64_t ptile = pitem / n_chunks;
                    const int64_t pcol0 = (int64_t)(pitem % n_chunks) * W;
                    const int pw = (int)((row_len - pcol0) < W ? (row_len - pcol0) : W);
                    const int pn = tile_nrows[ptile];
                    const int32_t* prow = tile_rows + ptile * cap;
                    for (int r = lane; r < pn; r += 32)
                        l2_prefetch_bulk(data + (int64_t)prow[r] * row_len + pcol0, (uint32_t)pw * 4u);
                }
            }
            if (G4) {
                // stage rows in quads; a partial last quad repeats padding rows (index 0), columns past the row end
                // are zero-filled by the TMA unit; every request delivers the full box (4 * W * 4 bytes)
                const int nquad = (nstage + 3) >> 2;
                if (lane == 0) {
                    mbar_expect_tx(&full_bar[s], (uint32_t)nquad * (uint32_t)(16 * W) + (uint32_t)cap * 6u);
                    tma_load_1d(s_w, w + tile * cap, (uint32_t)cap * 4u, &full_bar[s]);
                    tma_load_1d(s_lidx, tile_lidx + tile * cap, (uint32_t)cap * 2u, &full_bar[s]);
                }
                __syncwarp();
                for (int qd = lane; qd < nquad; qd += 32) {
                    const int4 rr = *reinterpret_cast<const int4*>(rows + 4 * qd);
                    tma_gather4(s_rows + (size_t)(4 * qd) * W, &tmap, (int)col0, rr.x, rr.y, rr.z, rr.w, &full_bar[s]);
                }
            } else {
                if (lane == 0) {
                    mbar_expect_tx(&full_bar[s], (uint32_t)nstage * row_bytes + (uint32_t)cap * 6u);
                    tma_load_1d(s_w, w + tile * cap, (uint32_t)cap * 4u, &full_bar[s]);
                    tma_load_1d(s_lidx, tile_lidx + tile * cap, (uint32_t)cap * 2u, &full_bar[s]);
                }
                __syncwarp();
                for (int r = lane; r < nstage; r += 32)
                    tma_load_1d(s_rows + (size_t)r * W, data + (int64_t)rows[r] * row_len + col0, row_bytes,
                                &full_bar[s]);
            }
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        constexpr int S = W / 128;
        int64_t it = 0;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int s = (int)(it % n_stages);
            const uint32_t round = (uint32_t)(it / n_stages);
            const int64_t tile = item / n_chunks;
            const int chunk = (int)(item % n_chunks);
            const int64_t cell0 = tile * kPipeTileCells;
            const int ncell = (int)((n_cells - cell0) < kPipeTileCells ? (n_cells - cell0) : kPipeTileCells);
            const int64_t col0 = (int64_t)chunk * W;
            const int wcur = (int)((row_len - col0) < W ? (row_len - col0) : W);
            const int nrows = tile_nrows[tile];
            const int nstage = nrows < r_smem ? nrows : r_smem;
            unsigned char* base = smem_raw + (size_t)s * stage_bytes;
            const float* s_rows = reinterpret_cast<const float*>(base);
            const float* s_w = reinterpret_cast<const float*>(base + (size_t)r_smem * W * 4);
            const uint16_t* s_lidx = reinterpret_cast<const uint16_t*>(s_w + cap);
            const int32_t* rows = tile_rows + tile * cap;
            mbar_wait(&full_bar[s], round & 1u);
            for (int c = warp; c < ncell; c += kPipeConsumerWarps) {
                float4 acc[S];
#pragma unroll
                for (int q = 0; q < S; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint16_t* li = s_lidx + c * k;
                const float* cw = s_w + c * k;
                for (int j = 0; j < k; ++j) {
                    const int r = li[j];
                    const float wj = cw[j];
                    if (r < nstage) {
                        const float4* src = reinterpret_cast<const float4*>(s_rows + (size_t)r * W) + lane;
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            const float4 x = src[q * 32];
                            acc[q].x = fmaf(wj, x.x, acc[q].x); acc[q].y = fmaf(wj, x.y, acc[q].y);
                            acc[q].z = fmaf(wj, x.z, acc[q].z); acc[q].w = fmaf(wj, x.w, acc[q].w);
                        }
                    } else {
                        const float* g = data + (int64_t)rows[r] * row_len + col0;
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            const int cc = q * 128 + lane * 4;
                            if (cc < wcur) {
                                const float4 x = *reinterpret_cast<const float4*>(g + cc);
                                acc[q].x = fmaf(wj, x.x, acc[q].x); acc[q].y = fmaf(wj, x.y, acc[q].y);
                                acc[q].z = fmaf(wj, x.z, acc[q].z); acc[q].w = fmaf(wj, x.w, acc[q].w);
                            }
                        }
                    }
                }
                const int64_t orow = out_row ? (int64_t)out_row[cell0 + c] : (cell0 + c);
                float* o = out + orow * row_len + col0;
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    const int cc = q * 128 + lane * 4;
                    if (cc < wcur) *reinterpret_cast<float4*>(o + cc) = acc[q];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
}

}  // namespace s3

using namespace s3;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int s3_interp_pipelined(const float* d_data, int64_t n_src, int64_t row_len, const int32_t* d_tile_rows,
                                   const int32_t* d_tile_nrows, const uint16_t* d_tile_lidx, const float* d_w,
                                   int64_t n_cells, int k, int max_rows, int chunk_cols, int stage_rows, int n_ctas,
                                   int use_gather4, const int32_t* d_out_row, float* d_out, void* stream) {
    S3_REQUIRE(d_data && d_tile_rows && d_tile_nrows && d_tile_lidx && d_w && d_out, "s3_interp_pipelined: NULL argument");
    S3_REQUIRE(row_len % 4 == 0 && ((uintptr_t)d_data % 16) == 0 && ((uintptr_t)d_out % 16) == 0,
               "s3_interp_pipelined: rows must be 16-byte aligned (row_len %% 4 == 0)");
    S3_REQUIRE(chunk_cols == 128 || chunk_cols == 256, "s3_interp_pipelined: chunk_cols must be 128 or 256");
    S3_REQUIRE(kPipeTileCells * k <= 2048 && (kPipeTileCells * k) % 8 == 0, "s3_interp_pipelined: unsupported k=%d", k);
    if (n_cells == 0 || row_len == 0) return S3_OK;
    const int cap = kPipeTileCells * k;
    int r_smem = stage_rows > 0 ? stage_rows : max_rows;
    if (r_smem > max_rows) r_smem = max_rows;
    if (r_smem < 1) r_smem = 1;
    const size_t budget = 220 * 1024;
    size_t stage_bytes = (((size_t)r_smem * chunk_cols * 4 + (size_t)cap * 6) + 127) & ~(size_t)127;
    if (stage_bytes * 2 > budget) {   // keep at least two stages
        r_smem = (int)((budget / 2 - (size_t)cap * 6 - 128) / ((size_t)chunk_cols * 4));
        stage_bytes = (((size_t)r_smem * chunk_cols * 4 + (size_t)cap * 6) + 127) & ~(size_t)127;
    }
    int n_stages = (int)(budget / stage_bytes);
    if (n_stages > kPipeMaxStages) n_stages = kPipeMaxStages;
    const size_t smem = stage_bytes * n_stages;
    const int n_chunks = (int)ceil_div(row_len, chunk_cols);
    const int64_t tiles = ceil_div(n_cells, kPipeTileCells);
    const int64_t n_items = tiles * n_chunks;
    int grid = n_ctas > 0 ? n_ctas : kNumSMs;
    if ((int64_t)grid > n_items) grid = (int)n_items;
    cudaStream_t st = (cudaStream_t)stream;
    const bool g4 = use_gather4 != 0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (g4) {
        r_smem &= ~3;   // whole quads
        S3_REQUIRE(r_smem >= 4, "s3_interp_pipelined: staging buffer too small for gather4");
        static PFN_encodeTiled encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            S3_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            S3_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
            encode = (PFN_encodeTiled)fn;
        }
        const cuuint64_t gdim[2] = {(cuuint64_t)row_len, (cuuint64_t)n_src};
        const cuuint64_t gstride[1] = {(cuuint64_t)row_len * 4};
        const cuuint32_t box[2] = {(cuuint32_t)chunk_cols, 1};
        const cuuint32_t estride[2] = {1, 1};
        CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_data), gdim, gstride, box,
                             estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        S3_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)cr);
    }
    stage_bytes = (((size_t)r_smem * chunk_cols * 4 + (size_t)cap * 6) + 127) & ~(size_t)127;
    n_stages = (int)(budget / stage_bytes);
    if (n_stages > kPipeMaxStages) n_stages = kPipeMaxStages;
    const size_t smem2 = stage_bytes * n_stages;
#define S3_LAUNCH_PIPE(WW, GG)                                                                                        \
    do {                                                                                                              \
        S3_CUDA(cudaFuncSetAttribute(interp_pipe_kernel<WW, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                     (int)smem2));                                                                    \
        interp_pipe_kernel<WW, GG><<<grid, kPipeThreads, smem2, st>>>(tmap, d_data, row_len, d_tile_rows, d_tile_nrows, \
                                                                      d_tile_lidx, d_w, n_cells, k, n_chunks, n_items, \
                                                                      r_smem, n_stages, g_pipe_prefetch, d_out_row, d_out); \
    } while (0)
    if (chunk_cols == 128 && g4) S3_LAUNCH_PIPE(128, true);
    else if (chunk_cols == 128) S3_LAUNCH_PIPE(128, false);
    else if (g4) S3_LAUNCH_PIPE(256, true);
    else S3_LAUNCH_PIPE(256, false);
#undef S3_LAUNCH_PIPE
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}
