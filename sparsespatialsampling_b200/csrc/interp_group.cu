// Grouped export interpolation: a warp interpolates G consecutive cells (Morton neighbours) at once.
//
//   out[c, :] = sum_j w[c, j] * data[idx[c, j], :]            (interpolate_data, sparseSpatialSampling/export.py:446-468)
//
// The warp-per-cell kernel (interp.cu) is bound by the global-load path of the SM (LDG ~64 B/clk/SM): every output value
// pulls k source values through L1, whatever the cache hit rate. Neighbouring cells share most of their k nearest
// points (C2: 4 consecutive cells reference 19.4 distinct rows, not 32), so here every row of the group's UNION is
// loaded ONCE into registers and folded into up to G accumulators -- 39 % fewer loads per output for k = 8, 36 % for
// k = 26 -- at the price of a small per-group table that is built once per grid (s3_interp_groups_build):
//   cnt[g]                      number of distinct source rows of group g
//   ent[g][u] = {row, mask}     u-th distinct row (order of first appearance = nearest first), bit c of mask set if
//                               cell c of the group uses it
//   wts[g][u][c]                its weight for cell c (0 where the mask bit is clear)
// Rows a cell does not use are skipped by predicate, not multiplied by zero, so a non-finite value in a neighbour's
// row cannot leak into a cell that does not reference it. Per cell the terms are accumulated in union order, i.e. a
// different (still fixed) fp32 summation order than the warp-per-cell kernel.
#include "common.cuh"
#include "../../include/s3b200.h"

namespace s3 {

constexpr int kGroup = 4;             // cells per warp; wts rows are float4
constexpr int kPad = 8;               // shared-memory lists are padded to a multiple of the batch size (<= 8)

// One thread per group; the lists are short (<= G * k entries) and the table is built once per grid.
__global__ void __launch_bounds__(128)
interp_groups_build_kernel(const int32_t* __restrict__ idx, const float* __restrict__ w, int64_t n_cells, int k,
                           int64_t n_groups, int32_t* __restrict__ cnt, int2* __restrict__ ent,
                           float4* __restrict__ wts) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int stride = kGroup * k;
    int2* e = ent + g * stride;
    float* wt = reinterpret_cast<float*>(wts + g * stride);
    int n = 0;
    for (int c = 0; c < kGroup; ++c) {
        const int64_t cell = g * kGroup + c;
        if (cell >= n_cells) break;
        for (int j = 0; j < k; ++j) {
            const int32_t r = idx[cell * k + j];
            const float wj = w[cell * k + j];
            int u = 0;
            while (u < n && e[u].x != r) ++u;
            if (u == n) {
                e[n] = make_int2(r, 0);
                wt[4 * n + 0] = 0.f; wt[4 * n + 1] = 0.f; wt[4 * n + 2] = 0.f; wt[4 * n + 3] = 0.f;
                ++n;
            }
            // a row listed twice for one cell (duplicate neighbour) cannot be expressed by one weight: sum them
            wt[4 * u + c] = (e[u].y >> c) & 1 ? wt[4 * u + c] + wj : wj;
            e[u].y |= 1 << c;
        }
    }
    cnt[g] = n;
}

// Ordinary (ordered) 128-bit global load as a volatile asm statement, see the scheduling fence in the kernel.
__device__ __forceinline__ float4 ldg_v4(const char* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// BATCH distinct rows are in flight per lane. Left alone, ptxas sinks every load next to its first use (one load in
// flight per warp: 0.82 ms on C2 instead of 0.48 with four), so the loop is fenced: the BATCH row loads are ordinary
// global loads issued BEFORE a warp barrier, the weights are read from shared memory AFTER it, and since memory
// operations do not cross the barrier no FFMA of the batch can be scheduled in front of a load.
// Shared memory per warp: the byte offsets of the group's distinct rows (row * row_len * 4, computed when the table is
// staged) and their float4 weights, padded to a multiple of BATCH with the last row at zero weights, so the loop has
// no tail. A cell that does not use a row has weight 0 there and its FFMAs are skipped by predicate (IDW weights are
// > 0, export.py:420-424), so a non-finite value cannot leak into it. Lanes past the end of the row load column col0
// and do not store.
template <int BATCH, int MINB, int UNROLL>
__global__ void __launch_bounds__(256, MINB)
interp_group_kernel(const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ cnt,
                    const int2* __restrict__ ent, const float4* __restrict__ wts, int64_t n_cells, int k,
                    const int32_t* __restrict__ out_row, float* __restrict__ out, int64_t chunk_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int V = 4;
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t group = (int64_t)blockIdx.x * warps + warp;
    const int64_t cell0 = group * kGroup;
    if (cell0 >= n_cells) return;
    const int stride = kGroup * k;
    const int sstride = stride + kPad;                   // shared-memory list: padded to a multiple of BATCH
    float4* s_w = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * sstride;
    int64_t* s_off = reinterpret_cast<int64_t*>(smem_raw + (size_t)warps * sstride * sizeof(float4)) + (size_t)warp * sstride;
    const int n = cnt[group];
    const int n_pad = ((n + BATCH - 1) / BATCH) * BATCH;
    for (int u = lane; u < n_pad; u += 32) {
        const int v = u < n ? u : n - 1;                 // padding: the last row again, with zero weights
        s_w[u] = u < n ? wts[group * stride + u] : make_float4(0.f, 0.f, 0.f, 0.f);
        s_off[u] = (int64_t)ent[group * stride + v].x * row_len * (int64_t)sizeof(float);
    }
    __syncwarp();
    const int64_t col_begin = (int64_t)blockIdx.y * chunk_cols;
    const int64_t col_end = (col_begin + chunk_cols) < row_len ? (col_begin + chunk_cols) : row_len;
    float* o[kGroup];
#pragma unroll
    for (int c = 0; c < kGroup; ++c) {
        const int64_t cell = cell0 + c;
        o[c] = cell < n_cells ? out + (out_row ? (int64_t)out_row[cell] : cell) * row_len + lane * V : nullptr;
    }
    constexpr int STEP = 32 * V;
    // UNROLL column vectors per lane and step: the warp reads UNROLL * 512 contiguous bytes of every row
    for (int64_t col0 = col_begin; col0 < col_end; col0 += STEP * UNROLL) {
        float acc[UNROLL][kGroup][V];
        bool in[UNROLL];
        int uoff[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            in[u] = col0 + u * STEP + lane * V < col_end;
            uoff[u] = in[u] ? u * STEP * (int)sizeof(float) : 0;
#pragma unroll
            for (int c = 0; c < kGroup; ++c)
#pragma unroll
                for (int e = 0; e < V; ++e) acc[u][c][e] = 0.f;
        }
        const char* colbase = reinterpret_cast<const char*>(data + (in[0] ? col0 + lane * V : col0));
        for (int j = 0; j < n_pad; j += BATCH) {
            float4 x[BATCH][UNROLL];
#pragma unroll
            for (int b = 0; b < BATCH; ++b)
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) x[b][u] = ldg_v4(colbase + s_off[j + b] + uoff[u]);
            __syncwarp();                                          // scheduling fence, see above
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const float4 wv = s_w[j + b];
                const float wc[kGroup] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int c = 0; c < kGroup; ++c) {
                    if (wc[c] != 0.f) {
#pragma unroll
                        for (int u = 0; u < UNROLL; ++u) {
                            const float xv[V] = {x[b][u].x, x[b][u].y, x[b][u].z, x[b][u].w};
#pragma unroll
                            for (int e = 0; e < V; ++e) acc[u][c][e] = fmaf(wc[c], xv[e], acc[u][c][e]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (in[u]) {
#pragma unroll
                for (int c = 0; c < kGroup; ++c) {
                    if (o[c] != nullptr)
                        *reinterpret_cast<float4*>(o[c] + col0 + u * STEP) =
                            make_float4(acc[u][c][0], acc[u][c][1], acc[u][c][2], acc[u][c][3]);
                }
            }
        }
    }
}

static int g_group_warps = 8;     // warps (= 4-cell groups) per CTA (s3_set_tuning key 15)
static int g_group_batch = 4;     // distinct rows in flight per lane (s3_set_tuning key 16)
static int g_group_unroll = 1;    // column vectors per lane and step (s3_set_tuning key 18)
static int g_group_minb = 4;      // CTAs of 256 threads per SM the register allocation must allow (s3_set_tuning key 17)

int set_group_tuning(int key, int value) {
    if (key == 15) {
        S3_REQUIRE(value >= 1 && value <= 8, "s3_set_tuning: warps per CTA of the grouped kernel must be 1..8");
        g_group_warps = value;
        return S3_OK;
    }
    if (key == 18) {
        S3_REQUIRE(value == 1 || value == 2, "s3_set_tuning: column vectors per lane of the grouped kernel must be 1 or 2");
        g_group_unroll = value;
        return S3_OK;
    }
    if (key == 17) {
        S3_REQUIRE(value >= 2 && value <= 6, "s3_set_tuning: CTAs per SM of the grouped kernel must be 2..6");
        g_group_minb = value;
        return S3_OK;
    }
    S3_REQUIRE(value == 1 || value == 2 || value == 3 || value == 4 || value == 6 || value == 8,
               "s3_set_tuning: rows in flight of the grouped kernel must be 1, 2, 3, 4, 6 or 8");
    g_group_batch = value;
    return S3_OK;
}

}  // namespace s3

using namespace s3;

extern "C" int s3_interp_group_size(void) { return kGroup; }

extern "C" int s3_interp_groups_build(const int32_t* d_idx, const float* d_w, int64_t n_cells, int k, int32_t* d_cnt,
                                      void* d_ent, void* d_wts, void* stream) {
    S3_REQUIRE(n_cells >= 0, "s3_interp_groups_build: bad sizes");
    if (n_cells == 0) return S3_OK;
    S3_REQUIRE(d_idx && d_w && d_cnt && d_ent && d_wts, "s3_interp_groups_build: NULL argument");
    S3_REQUIRE(k >= 1 && k <= 64, "s3_interp_groups_build: k=%d out of range", k);
    const int64_t n_groups = (n_cells + kGroup - 1) / kGroup;
    const int64_t blocks = (n_groups + 127) / 128;
    S3_REQUIRE(blocks < ((int64_t)1 << 31), "s3_interp_groups_build: too many cells");
    interp_groups_build_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        d_idx, d_w, n_cells, k, n_groups, d_cnt, reinterpret_cast<int2*>(d_ent), reinterpret_cast<float4*>(d_wts));
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}

extern "C" int s3_interp_grouped(const void* d_data, int64_t n_src, int64_t row_len, const int32_t* d_cnt,
                                 const void* d_ent, const void* d_wts, int64_t n_cells, int k,
                                 const int32_t* d_out_row, void* d_out, int out_dtype, void* stream) {
    S3_REQUIRE(n_cells >= 0 && row_len >= 0, "s3_interp_grouped: bad sizes");
    if (n_cells == 0 || row_len == 0) return S3_OK;
    S3_REQUIRE(d_data && d_cnt && d_ent && d_wts && d_out, "s3_interp_grouped: NULL argument");
    S3_REQUIRE(k >= 1 && k <= 64 && n_src >= 1, "s3_interp_grouped: k=%d out of range", k);
    S3_REQUIRE(row_len % 4 == 0 && ((uintptr_t)d_data) % 16 == 0 && ((uintptr_t)d_out) % 16 == 0,
               "s3_interp_grouped: rows must be multiples of 4 columns and 16-byte aligned (use s3_interp_gather)");
    S3_REQUIRE(out_dtype == S3_F32, "s3_interp_grouped: fp32 in, fp32 out only (dtype %d); use s3_interp_gather", out_dtype);
    const int warps = g_group_warps;
    const int64_t n_groups = (n_cells + kGroup - 1) / kGroup;
    const int64_t blocks = (n_groups + warps - 1) / warps;
    S3_REQUIRE(blocks < ((int64_t)1 << 31), "s3_interp_grouped: too many cells");
    const size_t smem = (size_t)warps * (kGroup * k + kPad) * (sizeof(float4) + sizeof(int64_t));
    const dim3 grid((unsigned)blocks, 1);
    const int64_t chunk = (int64_t)1 << 40;
    const float* data = reinterpret_cast<const float*>(d_data);
    const int2* ent = reinterpret_cast<const int2*>(d_ent);
    const float4* wts = reinterpret_cast<const float4*>(d_wts);
    cudaStream_t st = (cudaStream_t)stream;
#define S3_GROUP(BB, MM, UU)                                                                                      \
    do {                                                                                                          \
        if (smem > 48 * 1024)                                                                                     \
            S3_CUDA(cudaFuncSetAttribute(interp_group_kernel<BB, MM, UU>,                                           \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        interp_group_kernel<BB, MM, UU><<<grid, warps * 32, smem, st>>>(                                            \
            data, row_len, d_cnt, ent, wts, n_cells, k, d_out_row, reinterpret_cast<float*>(d_out), chunk);        \
    } while (0)
    const int code = g_group_unroll * 100 + g_group_batch * 10 + g_group_minb;
    switch (code) {
        case 124: S3_GROUP(2, 4, 1); break;
        case 125: S3_GROUP(2, 5, 1); break;
        case 145: S3_GROUP(4, 5, 1); break;
        case 143: S3_GROUP(4, 3, 1); break;
        case 163: S3_GROUP(6, 3, 1); break;
        case 164: S3_GROUP(6, 4, 1); break;
        case 183: S3_GROUP(8, 3, 1); break;
        case 182: S3_GROUP(8, 2, 1); break;
        case 222: S3_GROUP(2, 2, 2); break;
        case 223: S3_GROUP(2, 3, 2); break;
        case 224: S3_GROUP(2, 4, 2); break;
        case 242: S3_GROUP(4, 2, 2); break;
        case 243: S3_GROUP(4, 3, 2); break;
        case 232: S3_GROUP(3, 2, 2); break;
        case 233: S3_GROUP(3, 3, 2); break;
        default: S3_GROUP(4, 4, 1); break;
    }
#undef S3_GROUP
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}
