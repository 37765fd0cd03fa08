// Export-stage interpolation of snapshot rows onto the sampled grid:
//   out[c, :] = sum_j w[c, j] * data[idx[c, j], :]
// Replaces interpolate_data (sparseSpatialSampling/export.py:446-468), which materialises
// data[idx] ([chunk, k, D, T]) and reduces it on the CPU.
//
// Layout: data is [N, L] with L = D*T contiguous per source point (the reference's [N, D, T]),
// out is [Nc, L]. A CTA owns a tile of kCellsPerCta consecutive cells (callers pass cells in
// Morton order so neighbouring cells share source rows -> L1/L2 hits) times one column chunk
// of the row. The (idx, w) tile is staged into shared memory with one 1-D TMA bulk copy
// (cp.async.bulk + mbarrier); every lane then streams 128-bit column vectors of the k source
// rows (k independent loads in flight per thread) and writes one 128-bit result.
#include <type_traits>
#include "common.cuh"
#include "tma.cuh"
#include "../../include/s3b200.h"

namespace s3 {

extern int g_staging;
extern int g_stage_budget_kb;
extern int g_pipe_prefetch;
extern int g_tc_seg_kblocks;
extern int g_tc_flush_segments;
extern int g_tc_pair;
int set_group_tuning(int key, int value);
constexpr int kInterpThreads = 128;
constexpr int kMaxCellsPerCta = 32;
static int g_cells_per_cta = 4;
static int g_direct_variant = 1;   // 0 = CTA walks cells, 1 = warp per cell (s3_set_tuning key 3)
static int g_warps_per_cta = 8;
static int g_direct_regs = 0;      // k = 8 / 26: (idx, w) in registers instead of shuffle broadcasts (s3_set_tuning key 8)
static int g_bcast = -1;           // (idx, w) broadcast in the warp-per-cell kernels: 0 = SHFL, 1 = REDUX.OR, 2 / 3 = LDS.64 / LDS.128 from
                                   // shared memory, -1 = by k (k > 16: 2, else 0; measured) (s3_set_tuning key 13)
static int g_carveout = -1;        // shared-memory carve-out (percent) requested for the warp-per-cell kernel, -1 = driver default (s3_set_tuning key 20)
static int g_direct_window = 1;    // k > 16: window formulation of the warp-per-cell kernel (s3_set_tuning key 12)
static int g_chunk_cols = 0;       // columns per grid.y window of the warp-per-cell kernel, 0 = by k (s3_set_tuning key 9)
static int g_direct_sync = 0;      // barrier per column step in the warp-per-cell kernel (s3_set_tuning key 7)
static int g_unroll = 0;           // column vectors per lane and step (s3_set_tuning key 5)    // warp-per-cell variant (s3_set_tuning key 4)   // tuning knob (s3_set_tuning key 0)

template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ld_stream(const T* p) {
    // read-only path; rows are re-used by neighbouring cells, so let them allocate in L1
    return *reinterpret_cast<const Vec<T, V>*>(p);
}

// MODE 0: fp32 FMA accumulate (fast path). MODE 1: fp64, products then sequential adds without
// contraction -- the reference's (w * data[idx]).sum(dim=1) evaluation order.
template <typename Tin, typename Tw, typename Tout, int V, int MODE>
__global__ void __launch_bounds__(kInterpThreads)
interp_gather_kernel(const Tin* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                     const Tw* __restrict__ w, int64_t n_cells, int k, const int32_t* __restrict__ out_row,
                     Tout* __restrict__ out, int kCellsPerCta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* s_idx = reinterpret_cast<int32_t*>(smem_raw);
    Tw* s_w = reinterpret_cast<Tw*>(smem_raw + ((sizeof(int32_t) * kCellsPerCta * k + 15) / 16) * 16);
    __shared__ __align__(8) uint64_t bar;

    const int64_t cell0 = (int64_t)blockIdx.x * kCellsPerCta;
    const int ncell = (int)((n_cells - cell0) < kCellsPerCta ? (n_cells - cell0) : kCellsPerCta);
    const uint32_t bytes_idx = (uint32_t)(sizeof(int32_t) * ncell * k);
    const uint32_t bytes_w = (uint32_t)(sizeof(Tw) * ncell * k);
    const int32_t* g_idx = idx + cell0 * k;
    const Tw* g_w = w + cell0 * k;
    const bool bulk = ((bytes_idx | bytes_w) & 15u) == 0 && ((((uintptr_t)g_idx) | ((uintptr_t)g_w)) & 15u) == 0;

    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, bytes_idx + bytes_w);
            tma_load_1d(s_idx, g_idx, bytes_idx, &bar);
            tma_load_1d(s_w, g_w, bytes_w, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        for (int i = threadIdx.x; i < ncell * k; i += kInterpThreads) {
            s_idx[i] = g_idx[i];
            s_w[i] = g_w[i];
        }
        __syncthreads();
    }

    const int64_t col = ((int64_t)blockIdx.y * kInterpThreads + threadIdx.x) * V;
    if (col >= row_len) return;
    const Tin* dcol = data + col;

    for (int c = 0; c < ncell; ++c) {
        const int32_t* ci = s_idx + c * k;
        const Tw* cw = s_w + c * k;
        Tw acc[V];
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = (Tw)0;
        int j = 0;
        for (; j + 8 <= k; j += 8) {
            Vec<Tin, V> x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = ld_stream<Tin, V>(dcol + (int64_t)ci[j + u] * row_len);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const Tw wj = cw[j + u];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if (MODE == 0) acc[e] = fmaf((float)wj, (float)x[u].v[e], (float)acc[e]);
                    else acc[e] = __dadd_rn((double)acc[e], __dmul_rn((double)wj, (double)x[u].v[e]));
                }
            }
        }
        for (; j < k; ++j) {
            Vec<Tin, V> x = ld_stream<Tin, V>(dcol + (int64_t)ci[j] * row_len);
            const Tw wj = cw[j];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (MODE == 0) acc[e] = fmaf((float)wj, (float)x.v[e], (float)acc[e]);
                else acc[e] = __dadd_rn((double)acc[e], __dmul_rn((double)wj, (double)x.v[e]));
            }
        }
        const int64_t orow = out_row ? (int64_t)out_row[cell0 + c] : (cell0 + c);
        Vec<Tout, V> o;
#pragma unroll
        for (int e = 0; e < V; ++e) o.v[e] = (Tout)acc[e];
        *reinterpret_cast<Vec<Tout, V>*>(out + orow * row_len + col) = o;
    }
}

// Broadcast of lane `src`'s value to the whole warp. BCAST = 0: SHFL (one wavefront of the LSU / L1 data pipe -- the
// pipe that bounds the warp-per-cell kernel). BCAST = 1: REDUX.OR over (lane == src ? bits : 0): no L1 wavefront, the
// result lands in a uniform register.
template <int BCAST>
__device__ __forceinline__ uint32_t bcast_bits(uint32_t v, int src, int lane) {
    if (BCAST == 1) return __reduce_or_sync(0xffffffffu, lane == src ? v : 0u);
    return __shfl_sync(0xffffffffu, v, src);
}
template <int BCAST>
__device__ __forceinline__ int32_t bcast(int32_t v, int src, int lane) { return (int32_t)bcast_bits<BCAST>((uint32_t)v, src, lane); }
template <int BCAST>
__device__ __forceinline__ float bcast(float v, int src, int lane) { return __uint_as_float(bcast_bits<BCAST>(__float_as_uint(v), src, lane)); }
template <int BCAST>
__device__ __forceinline__ double bcast(double v, int src, int lane) {
    const uint32_t lo = bcast_bits<BCAST>((uint32_t)__double2loint(v), src, lane);
    const uint32_t hi = bcast_bits<BCAST>((uint32_t)__double2hiint(v), src, lane);
    return __hiloint2double((int)hi, (int)lo);
}

// Warp-per-cell variant: the warps of a CTA work on CONSECUTIVE cells (Morton neighbours) and sweep the row in
// lock step, 128 columns (one 128-bit vector per lane) at a time. Neighbouring cells share most of their source
// rows, so the same 512-byte row segments are requested by several warps of the CTA within a few hundred cycles
// and are served by L1 (hit or hit-under-miss) instead of crossing L2->SM once per reference; the CTAs in flight
// cover a compact window of the grid, so the second use of a row by a neighbouring CTA is an L2 hit.
// (idx, w) of the cell live in registers (lane j holds neighbour j) and are broadcast with shuffles.
// SYNC: the warps of the CTA additionally meet at a barrier after every column step, so that the row segments shared by
// neighbouring cells are requested within one step by all warps (L1 hit / hit-under-miss instead of a second fill).
// CHUNKED: blockIdx.y selects a window of `chunk_cols` columns (slow grid dimension): all cells sweep window 0, then
// window 1, ... so that the row segments a wave of CTAs touches (cells in flight x unique rows x window bytes) stay
// inside the L2 for long rows. The un-chunked instantiation is kept byte for byte: this kernel's speed depends on
// the load/FMA interleaving ptxas picks for the neighbour loop (an explicit 4-deep batching measured 25 % slower, a
// different loop bound 10 % slower).
template <typename Tin, typename Tw, typename Tout, int V, int MODE, int UNROLL, bool SYNC, bool CHUNKED, int BCAST = 0>
__global__ void __launch_bounds__(512)
interp_warpcell_kernel(const Tin* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                       const Tw* __restrict__ w, int64_t n_cells, int k, const int32_t* __restrict__ out_row,
                       Tout* __restrict__ out, int64_t row_stride, int64_t chunk_cols) {
    if (CHUNKED) {
        // from here on `row_len` is the width of this window; rows are `row_stride` apart
        const int64_t begin = (int64_t)blockIdx.y * chunk_cols;
        data += begin;
        out += begin;
        row_len = (row_len - begin) < chunk_cols ? (row_len - begin) : chunk_cols;
    }
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * warps + warp;
    const bool active = cell < n_cells;
    if (!SYNC && !active) return;
    int32_t idx_lo = 0, idx_hi = 0;
    Tw w_lo = (Tw)0, w_hi = (Tw)0;
    if (active && lane < k) { idx_lo = idx[cell * k + lane]; w_lo = w[cell * k + lane]; }
    if (active && lane + 32 < k) { idx_hi = idx[cell * k + lane + 32]; w_hi = w[cell * k + lane + 32]; }
    const int64_t orow = !active ? 0 : (out_row ? (int64_t)out_row[cell] : cell);
    Tout* o = out + orow * (CHUNKED ? row_stride : row_len);
    constexpr int STEP = 32 * V;
    // BCAST >= 2: the warp parks its (index, weight) pairs in shared memory; the neighbour loop reads them back with one
    // uniform LDS.64 (BCAST 2) or one LDS.128 per two neighbours (BCAST 3) instead of two SHFL per neighbour.
    extern __shared__ int2 s_pairs[];
    const int pair_stride = (k + 1) & ~1;
    int2* my_pairs = s_pairs + warp * pair_stride;
    if (BCAST >= 2) {
        if (lane < k) my_pairs[lane] = make_int2(idx_lo, __float_as_int((float)w_lo));
        if (lane + 32 < k) my_pairs[lane + 32] = make_int2(idx_hi, __float_as_int((float)w_hi));
        if (lane == 0 && (k & 1)) my_pairs[k] = make_int2(0, 0);
        __syncwarp();
    }
    for (int64_t col0 = 0; col0 < row_len; col0 += (int64_t)STEP * UNROLL) {
        if (SYNC) {
            __syncthreads();
            if (!active) continue;
        }
        Tw acc[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[u][e] = (Tw)0;
        // One neighbour at a time on purpose: ncu shows this kernel bound by the L1 data pipe (LDG.128 = 4 wavefronts
        // at ~2 cycles each), not by latency; batching 8 neighbours' loads per lane (tried) only lowered the L1 hit
        // rate and cost 25 %.  The UNROLL column vectors of one neighbour are independent loads.
        if (BCAST == 3) {
            // two neighbours per LDS.128; a zero-weight pad entry (row 0) completes an odd k
            const int4* my_quads = reinterpret_cast<const int4*>(my_pairs);
            for (int j = 0; j < k; j += 2) {
                const int4 q = my_quads[j >> 1];
                const Tin* src0 = data + (int64_t)q.x * (CHUNKED ? row_stride : row_len) + col0 + lane * V;
                const Tin* src1 = data + (int64_t)q.z * (CHUNKED ? row_stride : row_len) + col0 + lane * V;
                const Tw w0 = (Tw)__int_as_float(q.y), w1 = (Tw)__int_as_float(q.w);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (col0 + u * STEP + lane * V < row_len) {
                        const Vec<Tin, V> x0 = ld_stream<Tin, V>(src0 + u * STEP);
                        const Vec<Tin, V> x1 = ld_stream<Tin, V>(src1 + u * STEP);
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            if (MODE == 0) {
                                acc[u][e] = fmaf((float)w0, (float)x0.v[e], (float)acc[u][e]);
                                if (j + 1 < k) acc[u][e] = fmaf((float)w1, (float)x1.v[e], (float)acc[u][e]);
                            } else {
                                acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)w0, (double)x0.v[e]));
                                if (j + 1 < k) acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)w1, (double)x1.v[e]));
                            }
                        }
                    }
                }
            }
        } else
        for (int j = 0; j < k; ++j) {
            int32_t r;
            Tw wj;
            if (BCAST == 2) {
                const int2 pr = my_pairs[j];
                r = pr.x;
                wj = (Tw)__int_as_float(pr.y);
            } else {
                r = BCAST ? bcast<BCAST>((j & 32) ? idx_hi : idx_lo, j & 31, lane)
                          : __shfl_sync(0xffffffffu, (j & 32) ? idx_hi : idx_lo, j & 31);
                wj = BCAST ? bcast<BCAST>((j & 32) ? w_hi : w_lo, j & 31, lane)
                           : __shfl_sync(0xffffffffu, (j & 32) ? w_hi : w_lo, j & 31);
            }
            const Tin* src = data + (int64_t)r * (CHUNKED ? row_stride : row_len) + col0 + lane * V;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (col0 + u * STEP + lane * V < row_len) {
                    const Vec<Tin, V> x = ld_stream<Tin, V>(src + u * STEP);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (MODE == 0) acc[u][e] = fmaf((float)wj, (float)x.v[e], (float)acc[u][e]);
                        else acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)wj, (double)x.v[e]));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t c = col0 + u * STEP + lane * V;
            if (c < row_len) {
                Vec<Tout, V> ov;
#pragma unroll
                for (int e = 0; e < V; ++e) ov.v[e] = (Tout)acc[u][e];
                *reinterpret_cast<Vec<Tout, V>*>(o + c) = ov;
            }
        }
    }
}

// Window formulation of the same kernel for the 3-D neighbour count (k = 26): the column range [col_begin, col_end) of
// blockIdx.y is swept with absolute column indices. Functionally identical to the CHUNKED instantiation above; it exists
// because ptxas schedules its neighbour loop differently (more loads in flight per warp), which measured 8-10 % faster
// for k = 26 with two column vectors per lane (C4: 7.0 ms vs 7.8 ms) and 10 % slower for k = 8 with one.
template <typename Tin, typename Tw, typename Tout, int V, int MODE, int UNROLL, int BCAST = 0>
__global__ void __launch_bounds__(512)
interp_warpcell_window_kernel(const Tin* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                              const Tw* __restrict__ w, int64_t n_cells, int k, const int32_t* __restrict__ out_row,
                              Tout* __restrict__ out, int64_t chunk_cols) {
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * warps + warp;
    if (cell >= n_cells) return;
    const int64_t col_begin = (int64_t)blockIdx.y * chunk_cols;
    const int64_t col_end = (col_begin + chunk_cols) < row_len ? (col_begin + chunk_cols) : row_len;
    int32_t idx_lo = 0, idx_hi = 0;
    Tw w_lo = (Tw)0, w_hi = (Tw)0;
    if (lane < k) { idx_lo = idx[cell * k + lane]; w_lo = w[cell * k + lane]; }
    if (lane + 32 < k) { idx_hi = idx[cell * k + lane + 32]; w_hi = w[cell * k + lane + 32]; }
    const int64_t orow = out_row ? (int64_t)out_row[cell] : cell;
    Tout* o = out + orow * row_len;
    constexpr int STEP = 32 * V;
    extern __shared__ int2 s_pairs[];                      // BCAST >= 2, see interp_warpcell_kernel
    const int pair_stride = (k + 1) & ~1;
    int2* my_pairs = s_pairs + warp * pair_stride;
    if (BCAST >= 2) {
        if (lane < k) my_pairs[lane] = make_int2(idx_lo, __float_as_int((float)w_lo));
        if (lane + 32 < k) my_pairs[lane + 32] = make_int2(idx_hi, __float_as_int((float)w_hi));
        if (lane == 0 && (k & 1)) my_pairs[k] = make_int2(0, 0);
        __syncwarp();
    }
    for (int64_t col0 = col_begin; col0 < col_end; col0 += (int64_t)STEP * UNROLL) {
        Tw acc[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[u][e] = (Tw)0;
        if (BCAST == 3) {
            const int4* my_quads = reinterpret_cast<const int4*>(my_pairs);
            for (int j = 0; j < k; j += 2) {
                const int4 q = my_quads[j >> 1];
                const Tin* src0 = data + (int64_t)q.x * row_len + col0 + lane * V;
                const Tin* src1 = data + (int64_t)q.z * row_len + col0 + lane * V;
                const Tw w0 = (Tw)__int_as_float(q.y), w1 = (Tw)__int_as_float(q.w);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (col0 + u * STEP + lane * V < col_end) {
                        const Vec<Tin, V> x0 = ld_stream<Tin, V>(src0 + u * STEP);
                        const Vec<Tin, V> x1 = ld_stream<Tin, V>(src1 + u * STEP);
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            if (MODE == 0) {
                                acc[u][e] = fmaf((float)w0, (float)x0.v[e], (float)acc[u][e]);
                                if (j + 1 < k) acc[u][e] = fmaf((float)w1, (float)x1.v[e], (float)acc[u][e]);
                            } else {
                                acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)w0, (double)x0.v[e]));
                                if (j + 1 < k) acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)w1, (double)x1.v[e]));
                            }
                        }
                    }
                }
            }
        } else
        for (int j = 0; j < k; ++j) {
            int32_t r;
            Tw wj;
            if (BCAST == 2) {
                const int2 pr = my_pairs[j];
                r = pr.x;
                wj = (Tw)__int_as_float(pr.y);
            } else {
                r = BCAST ? bcast<BCAST>((j & 32) ? idx_hi : idx_lo, j & 31, lane)
                          : __shfl_sync(0xffffffffu, (j & 32) ? idx_hi : idx_lo, j & 31);
                wj = BCAST ? bcast<BCAST>((j & 32) ? w_hi : w_lo, j & 31, lane)
                           : __shfl_sync(0xffffffffu, (j & 32) ? w_hi : w_lo, j & 31);
            }
            const Tin* src = data + (int64_t)r * row_len + col0 + lane * V;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (col0 + u * STEP + lane * V < col_end) {
                    const Vec<Tin, V> x = ld_stream<Tin, V>(src + u * STEP);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (MODE == 0) acc[u][e] = fmaf((float)wj, (float)x.v[e], (float)acc[u][e]);
                        else acc[u][e] = __dadd_rn((double)acc[u][e], __dmul_rn((double)wj, (double)x.v[e]));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t c = col0 + u * STEP + lane * V;
            if (c < col_end) {
                Vec<Tout, V> ov;
#pragma unroll
                for (int e = 0; e < V; ++e) ov.v[e] = (Tout)acc[u][e];
                *reinterpret_cast<Vec<Tout, V>*>(o + c) = ov;
            }
        }
    }
}

// Offset-table variant (fp32 in, fp32 weights): per neighbour the generic kernels spend ~24 instructions, most of them
// on the broadcast and on the 64-bit address r * row_len * 4 + base. Here the warp computes the BYTE OFFSET of each of
// its k source rows once per cell, parks {offset lo, offset hi, weight} in shared memory (16 B per neighbour) and the
// neighbour loop is LDS.128 (uniform address, one wavefront) -> 64-bit add -> LDG.128 -> 4 FFMA. Column windows as in
// interp_warpcell_window_kernel (blockIdx.y).
template <typename Tout, int V, int UNROLL>
__global__ void __launch_bounds__(512)
interp_warpcell_off_kernel(const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                           const float* __restrict__ w, int64_t n_cells, int k, const int32_t* __restrict__ out_row,
                           Tout* __restrict__ out, int64_t chunk_cols) {
    extern __shared__ int4 s_quads[];
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * warps + warp;
    if (cell >= n_cells) return;
    const int64_t col_begin = (int64_t)blockIdx.y * chunk_cols;
    const int64_t col_end = (col_begin + chunk_cols) < row_len ? (col_begin + chunk_cols) : row_len;
    int4* my = s_quads + warp * k;
    for (int j = lane; j < k; j += 32) {
        const int64_t ob = (int64_t)idx[cell * k + j] * row_len * (int64_t)sizeof(float);
        my[j] = make_int4((int)(uint32_t)ob, (int)(ob >> 32), __float_as_int(w[cell * k + j]), 0);
    }
    __syncwarp();
    const int64_t orow = out_row ? (int64_t)out_row[cell] : cell;
    Tout* o = out + orow * row_len;
    constexpr int STEP = 32 * V;
    for (int64_t col0 = col_begin; col0 < col_end; col0 += (int64_t)STEP * UNROLL) {
        float acc[UNROLL][V];
        bool in[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            in[u] = col0 + u * STEP + lane * V < col_end;
#pragma unroll
            for (int e = 0; e < V; ++e) acc[u][e] = 0.f;
        }
        const char* colbase = reinterpret_cast<const char*>(data + col0 + lane * V);
        for (int j = 0; j < k; ++j) {
            const int4 q = my[j];
            const int64_t ob = (int64_t)(((uint64_t)(uint32_t)q.y << 32) | (uint32_t)q.x);
            const float wj = __int_as_float(q.z);
            const float* src = reinterpret_cast<const float*>(colbase + ob);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (in[u]) {
                    const Vec<float, V> x = ld_stream<float, V>(src + u * STEP);
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[u][e] = fmaf(wj, x.v[e], acc[u][e]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (in[u]) {
                Vec<Tout, V> ov;
#pragma unroll
                for (int e = 0; e < V; ++e) ov.v[e] = (Tout)acc[u][e];
                *reinterpret_cast<Vec<Tout, V>*>(o + col0 + u * STEP + lane * V) = ov;
            }
        }
    }
}

// Fenced-batch variant of the offset-table kernel: BATCH neighbour rows are loaded (ordinary global loads) BEFORE a warp
// barrier and the weights are read from shared memory AFTER it, which pins the memory-level parallelism per warp to
// BATCH whatever ptxas would schedule (it otherwise sinks every load next to its first FFMA, interp_group.cu). The
// table is padded to a multiple of BATCH with the first row at weight 0.
template <typename Tout, int V, int BATCH>
__global__ void __launch_bounds__(512)
interp_warpcell_batch_kernel(const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                             const float* __restrict__ w, int64_t n_cells, int k, const int32_t* __restrict__ out_row,
                             Tout* __restrict__ out, int64_t chunk_cols) {
    extern __shared__ __align__(16) unsigned char bt_smem[];
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * warps + warp;
    if (cell >= n_cells) return;
    const int k_pad = ((k + BATCH - 1) / BATCH) * BATCH;
    int64_t* s_off = reinterpret_cast<int64_t*>(bt_smem) + (size_t)warp * k_pad;
    float* s_w = reinterpret_cast<float*>(bt_smem + (size_t)warps * k_pad * sizeof(int64_t)) + (size_t)warp * k_pad;
    for (int j = lane; j < k_pad; j += 32) {
        const int jj = j < k ? j : 0;
        s_off[j] = (int64_t)idx[cell * k + jj] * row_len * (int64_t)sizeof(float);
        s_w[j] = j < k ? w[cell * k + j] : 0.f;
    }
    __syncwarp();
    const int64_t col_begin = (int64_t)blockIdx.y * chunk_cols;
    const int64_t col_end = (col_begin + chunk_cols) < row_len ? (col_begin + chunk_cols) : row_len;
    const int64_t orow = out_row ? (int64_t)out_row[cell] : cell;
    Tout* o = out + orow * row_len;
    constexpr int STEP = 32 * V;
    for (int64_t col0 = col_begin; col0 < col_end; col0 += STEP) {
        float acc[V];
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = 0.f;
        const bool in = col0 + lane * V < col_end;
        const char* colbase = reinterpret_cast<const char*>(data + (in ? col0 + lane * V : col0));
        for (int j = 0; j < k_pad; j += BATCH) {
            float4 x[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(x[b].x), "=f"(x[b].y), "=f"(x[b].z), "=f"(x[b].w)
                             : "l"(colbase + s_off[j + b])
                             : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const float wj = s_w[j + b];
                acc[0] = fmaf(wj, x[b].x, acc[0]);
                acc[1] = fmaf(wj, x[b].y, acc[1]);
                acc[2] = fmaf(wj, x[b].z, acc[2]);
                acc[3] = fmaf(wj, x[b].w, acc[3]);
            }
        }
        if (in) {
            Vec<Tout, V> ov;
#pragma unroll
            for (int e = 0; e < V; ++e) ov.v[e] = (Tout)acc[e];
            *reinterpret_cast<Vec<Tout, V>*>(o + col0 + lane * V) = ov;
        }
    }
}

// Register-resident variant for the two neighbour counts S^3 uses (k = 8 in 2-D, 26 in 3-D; s_cube.py:161,
// export.py:117-118): every lane keeps the cell's k (index, weight) pairs in registers (uniform loads, one wavefront
// each), so the inner loop is address arithmetic + LDG.128 + FFMA only. The shuffle-broadcast of the generic kernel
// costs two LSU-pipe wavefronts per neighbour and column step -- a third of the wavefronts of a kernel that ncu shows
// bound by exactly that pipe.
template <int V, int UNROLL, int K>
__global__ void __launch_bounds__(512)
interp_warpcell_reg_kernel(const float* __restrict__ data, int64_t row_len, const int32_t* __restrict__ idx,
                           const float* __restrict__ w, int64_t n_cells, const int32_t* __restrict__ out_row,
                           float* __restrict__ out) {
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cell = (int64_t)blockIdx.x * warps + warp;
    if (cell >= n_cells) return;
    int32_t ri[K];
    float rw[K];
    {
        // K * 4 bytes per cell: 8-byte aligned for every even K
        const int2* pi = reinterpret_cast<const int2*>(idx + cell * K);
        const float2* pw = reinterpret_cast<const float2*>(w + cell * K);
#pragma unroll
        for (int j = 0; j < K / 2; ++j) {
            const int2 a = pi[j];
            const float2 b = pw[j];
            ri[2 * j] = a.x; ri[2 * j + 1] = a.y;
            rw[2 * j] = b.x; rw[2 * j + 1] = b.y;
        }
    }
    const int64_t orow = out_row ? (int64_t)out_row[cell] : cell;
    float* o = out + orow * row_len;
    constexpr int STEP = 32 * V;
    const float* base = data + lane * V;
    for (int64_t col0 = 0; col0 < row_len; col0 += (int64_t)STEP * UNROLL) {
        float acc[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[u][e] = 0.f;
        bool in[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) in[u] = col0 + u * STEP + lane * V < row_len;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float* src = base + (int64_t)ri[j] * row_len + col0;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (in[u]) {
                    const Vec<float, V> x = ld_stream<float, V>(src + u * STEP);
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[u][e] = fmaf(rw[j], x.v[e], acc[u][e]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (in[u]) {
                Vec<float, V> ov;
#pragma unroll
                for (int e = 0; e < V; ++e) ov.v[e] = acc[u][e];
                *reinterpret_cast<Vec<float, V>*>(o + col0 + u * STEP + lane * V) = ov;
            }
        }
    }
}

template <typename Tin, typename Tw, typename Tout, int MODE>
static int launch_interp(const void* data, int64_t row_len, const int32_t* idx, const void* w, int64_t n_cells, int k,
                         const int32_t* out_row, void* out, cudaStream_t stream) {
    if (n_cells == 0 || row_len == 0) return S3_OK;
    constexpr int VFULL = 16 / sizeof(Tin);
    const int kCellsPerCta = g_cells_per_cta;
    const bool vec_ok = (row_len % VFULL == 0) && (((uintptr_t)data) % 16 == 0) &&
                        (((uintptr_t)out) % (sizeof(Tout) * VFULL) == 0);
    const size_t smem = ((sizeof(int32_t) * kCellsPerCta * k + 15) / 16) * 16 + sizeof(Tw) * kCellsPerCta * k;
    const Tw* w_p = reinterpret_cast<const Tw*>(w);
    Tout* out_p = reinterpret_cast<Tout*>(out);
    if (g_direct_variant == 1 && k <= 64) {
        const int warps = g_warps_per_cta;
        const int64_t blocks = ceil_div(n_cells, warps);
        // column vectors per lane and step: measured best 1 for k = 8 (C2/C3), 2 for k = 26 (C4/C5); 0 = this rule
        const int unroll = g_unroll != 0 ? g_unroll : (k > 16 ? 2 : 1);
        S3_REQUIRE(blocks < ((int64_t)1 << 31), "too many cells");
        if (std::is_same<Tin, float>::value && std::is_same<Tout, float>::value && MODE == 0 && vec_ok &&
            g_direct_regs != 0 && (k == 8 || k == 26)) {
            const float* d32 = reinterpret_cast<const float*>(data);
            const float* w32 = reinterpret_cast<const float*>(w);
            float* o32 = reinterpret_cast<float*>(out);
#define S3_WARPCELL_REG(UU, KK)                                                                                 \
    interp_warpcell_reg_kernel<4, UU, KK><<<(unsigned)blocks, warps * 32, 0, stream>>>(d32, row_len, idx, w32,     \
                                                                                         n_cells, out_row, o32)
            if (k == 8) {
                if (unroll == 2) S3_WARPCELL_REG(2, 8); else S3_WARPCELL_REG(1, 8);
            } else {
                if (unroll == 2) S3_WARPCELL_REG(2, 26); else S3_WARPCELL_REG(1, 26);
            }
#undef S3_WARPCELL_REG
            S3_LAUNCH_CHECK();
            note_launch(1);
            return S3_OK;
        }
        // column windows (multiple of 256 columns); 0 = none. Measured on C4 (T = 2000, k = 26): windows of 512 / 1024 /
        // none: 7.32 / 7.41 / 7.27 ms -- the L2 already holds the working set, so windows are off by default
        int64_t chunk = g_chunk_cols > 0 ? g_chunk_cols : ((int64_t)1 << 40);
        chunk = ceil_div(chunk, 256) * 256;
        int64_t n_chunks = ceil_div(row_len, chunk);
        if (n_chunks > 65535) { chunk = ceil_div(ceil_div(row_len, 65535), 256) * 256; n_chunks = ceil_div(row_len, chunk); }
        const dim3 wc_grid((unsigned)blocks, (unsigned)n_chunks);
#define S3_WARPCELL(VV, UU, SS, CC)                                                                             \
    do {                                                                                                        \
        if (g_carveout >= 0)                                                                                    \
            cudaFuncSetAttribute(interp_warpcell_kernel<Tin, Tw, Tout, VV, MODE, UU, SS, CC>,                     \
                                 cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);                   \
        interp_warpcell_kernel<Tin, Tw, Tout, VV, MODE, UU, SS, CC><<<wc_grid, warps * 32, 0, stream>>>(           \
            reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, row_len, chunk);  \
    } while (0)
        const bool sync = g_direct_sync != 0;
        const int g_bcast = s3::g_bcast >= 0 ? s3::g_bcast : (k > 16 ? 2 : 0);      // shadows the knob: resolved per call
        if (g_bcast >= 52 && vec_ok && !sync && MODE == 0 && std::is_same<Tin, float>::value && std::is_same<Tw, float>::value) {
            const int batch = g_bcast - 50;
            const int k_pad = ((k + batch - 1) / batch) * batch;
            const size_t bsmem = (size_t)warps * k_pad * (sizeof(int64_t) + sizeof(float));
            const float* d32 = reinterpret_cast<const float*>(data);
            const float* w32 = reinterpret_cast<const float*>(w);
#define S3_BATCH(BB)                                                                                             \
    interp_warpcell_batch_kernel<Tout, 4, BB><<<wc_grid, warps * 32, bsmem, stream>>>(d32, row_len, idx, w32, n_cells, \
                                                                                        k, out_row, out_p, chunk)
            if (batch == 2) S3_BATCH(2); else if (batch == 8) S3_BATCH(8); else if (batch == 3) S3_BATCH(3); else S3_BATCH(4);
#undef S3_BATCH
        } else if (g_bcast == 4 && vec_ok && !sync && MODE == 0 && std::is_same<Tin, float>::value && std::is_same<Tw, float>::value) {
            const size_t off_smem = (size_t)warps * k * sizeof(int4);
            const float* d32 = reinterpret_cast<const float*>(data);
            const float* w32 = reinterpret_cast<const float*>(w);
            if (unroll == 2)
                interp_warpcell_off_kernel<Tout, 4, 2><<<wc_grid, warps * 32, off_smem, stream>>>(d32, row_len, idx, w32, n_cells, k, out_row, out_p, chunk);
            else
                interp_warpcell_off_kernel<Tout, 4, 1><<<wc_grid, warps * 32, off_smem, stream>>>(d32, row_len, idx, w32, n_cells, k, out_row, out_p, chunk);
        } else if (k > 16 && vec_ok && unroll == 2 && !sync && g_direct_window != 0) {
            const size_t pairs_smem = (size_t)warps * ((k + 1) & ~1) * sizeof(int2);
            if (g_bcast == 2 && std::is_same<Tw, float>::value)
                interp_warpcell_window_kernel<Tin, Tw, Tout, VFULL, MODE, 2, 2><<<wc_grid, warps * 32, pairs_smem, stream>>>(
                    reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, chunk);
            else if (g_bcast == 3 && std::is_same<Tw, float>::value)
                interp_warpcell_window_kernel<Tin, Tw, Tout, VFULL, MODE, 2, 3><<<wc_grid, warps * 32, pairs_smem, stream>>>(
                    reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, chunk);
            else if (g_bcast == 1)
                interp_warpcell_window_kernel<Tin, Tw, Tout, VFULL, MODE, 2, 1><<<wc_grid, warps * 32, 0, stream>>>(
                    reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, chunk);
            else
            interp_warpcell_window_kernel<Tin, Tw, Tout, VFULL, MODE, 2><<<wc_grid, warps * 32, 0, stream>>>(
                reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, chunk);
        } else if (g_bcast >= 2 && n_chunks == 1 && vec_ok && !sync && std::is_same<Tw, float>::value) {
            const size_t pairs_smem = (size_t)warps * ((k + 1) & ~1) * sizeof(int2);
#define S3_WARPCELL_LDS(UU, BB)                                                                                  \
    interp_warpcell_kernel<Tin, Tw, Tout, VFULL, MODE, UU, false, false, BB><<<wc_grid, warps * 32, pairs_smem, stream>>>( \
        reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, row_len, chunk)
            if (g_bcast == 2) { if (unroll == 2) S3_WARPCELL_LDS(2, 2); else S3_WARPCELL_LDS(1, 2); }
            else { if (unroll == 2) S3_WARPCELL_LDS(2, 3); else S3_WARPCELL_LDS(1, 3); }
#undef S3_WARPCELL_LDS
        } else if (g_bcast == 1 && n_chunks == 1 && vec_ok && !sync) {
            if (unroll == 2)
                interp_warpcell_kernel<Tin, Tw, Tout, VFULL, MODE, 2, false, false, 1><<<wc_grid, warps * 32, 0, stream>>>(
                    reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, row_len, chunk);
            else
                interp_warpcell_kernel<Tin, Tw, Tout, VFULL, MODE, 1, false, false, 1><<<wc_grid, warps * 32, 0, stream>>>(
                    reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, row_len, chunk);
        } else if (n_chunks > 1 && vec_ok) {
            if (unroll == 2) S3_WARPCELL(VFULL, 2, false, true); else S3_WARPCELL(VFULL, 1, false, true);
        } else if (vec_ok && unroll == 2) {
            if (sync) S3_WARPCELL(VFULL, 2, true, false); else S3_WARPCELL(VFULL, 2, false, false);
        } else if (vec_ok) {
            if (sync) S3_WARPCELL(VFULL, 1, true, false); else S3_WARPCELL(VFULL, 1, false, false);
        } else {
            S3_WARPCELL(1, 2, false, false);
        }
#undef S3_WARPCELL
        S3_LAUNCH_CHECK();
        note_launch(1);
        return S3_OK;
    }
    const int64_t tiles = ceil_div(n_cells, kCellsPerCta);
    S3_REQUIRE(tiles < ((int64_t)1 << 31), "too many cells");
    // blockIdx.x = cell tile (consecutive CTAs work on neighbouring cells), blockIdx.y = column chunk
    if (vec_ok) {
        dim3 grid((unsigned)tiles, (unsigned)ceil_div(row_len, (int64_t)kInterpThreads * VFULL));
        interp_gather_kernel<Tin, Tw, Tout, VFULL, MODE><<<grid, kInterpThreads, smem, stream>>>(
            reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, kCellsPerCta);
    } else {
        dim3 grid((unsigned)tiles, (unsigned)ceil_div(row_len, (int64_t)kInterpThreads));
        interp_gather_kernel<Tin, Tw, Tout, 1, MODE><<<grid, kInterpThreads, smem, stream>>>(
            reinterpret_cast<const Tin*>(data), row_len, idx, w_p, n_cells, k, out_row, out_p, kCellsPerCta);
    }
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}

}  // namespace s3

using namespace s3;

extern "C" int s3_set_tuning(int key, int value) {
    if (key == 0) {
        S3_REQUIRE(value >= 1 && value <= kMaxCellsPerCta, "s3_set_tuning: cells per CTA must be in [1, %d]", kMaxCellsPerCta);
        s3::g_cells_per_cta = value;
        return S3_OK;
    }
    if (key == 1) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: staging must be 0 (TMA) or 1 (cp.async)");
        s3::g_staging = value;
        return S3_OK;
    }
    if (key == 3) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: direct variant must be 0 or 1");
        s3::g_direct_variant = value;
        return S3_OK;
    }
    if (key == 4) {
        S3_REQUIRE(value >= 1 && value <= 16, "s3_set_tuning: warps per CTA must be 1..16");
        s3::g_warps_per_cta = value;
        return S3_OK;
    }
    if (key == 6) {
        S3_REQUIRE(value >= 0 && value <= 64, "s3_set_tuning: prefetch distance must be 0..64");
        s3::g_pipe_prefetch = value;
        return S3_OK;
    }
    if (key == 5) {
        S3_REQUIRE(value >= 0 && value <= 2, "s3_set_tuning: unroll must be 0 (auto), 1 or 2");
        s3::g_unroll = value;
        return S3_OK;
    }
    if (key == 2) {
        S3_REQUIRE(value >= 8 && value <= 200, "s3_set_tuning: staging budget must be 8..200 KB");
        s3::g_stage_budget_kb = value;
        return S3_OK;
    }
    if (key == 7) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: sync must be 0 or 1");
        s3::g_direct_sync = value;
        return S3_OK;
    }
    if (key == 8) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: register variant must be 0 or 1");
        s3::g_direct_regs = value;
        return S3_OK;
    }
    if (key == 9) {
        S3_REQUIRE(value >= 0, "s3_set_tuning: window columns must be >= 0");
        s3::g_chunk_cols = value;
        return S3_OK;
    }
    if (key >= 15 && key <= 18) return s3::set_group_tuning(key, value);
    if (key == 20) {
        S3_REQUIRE(value >= -1 && value <= 100, "s3_set_tuning: carve-out must be -1 (default) or 0..100 percent");
        s3::g_carveout = value;
        return S3_OK;
    }
    if (key == 13) {
        S3_REQUIRE((value >= -1 && value <= 4) || value == 52 || value == 53 || value == 54 || value == 58,
                   "s3_set_tuning: broadcast must be -1 (by k), 0 (SHFL), 1 (REDUX), 2 (LDS.64), 3 (LDS.128), 4 (offset table) or "
                   "52 / 53 / 54 / 58 (fenced batches of 2 / 3 / 4 / 8 rows)");
        s3::g_bcast = value;
        return S3_OK;
    }
    if (key == 12) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: window formulation must be 0 or 1");
        s3::g_direct_window = value;
        return S3_OK;
    }
    if (key == 10) {
        S3_REQUIRE(value >= 1 && value <= (1 << 20), "s3_set_tuning: K-blocks per TMEM segment must be >= 1");
        s3::g_tc_seg_kblocks = value;
        return S3_OK;
    }
    if (key == 14) {
        S3_REQUIRE(value == 0 || value == 1, "s3_set_tuning: paired Gram kernel must be 0 or 1");
        s3::g_tc_pair = value;
        return S3_OK;
    }
    if (key == 11) {
        S3_REQUIRE(value >= 1 && value <= (1 << 20), "s3_set_tuning: segments per fp64 flush must be >= 1");
        s3::g_tc_flush_segments = value;
        return S3_OK;
    }
    s3::set_error("s3_set_tuning: unknown key %d", key);
    return S3_ERR_INVALID;
}

extern "C" int s3_interp_gather(const void* d_data, int data_dtype, int64_t n_src, int64_t row_len,
                                const int32_t* d_idx, const void* d_w, int64_t n_cells, int k,
                                const int32_t* d_out_row, void* d_out, int out_dtype, void* stream) {
    S3_REQUIRE(n_cells >= 0 && row_len >= 0, "s3_interp_gather: bad sizes");
    if (n_cells == 0 || row_len == 0) return S3_OK;          // empty grids / empty rows: nothing to do (buffers may be NULL)
    S3_REQUIRE(d_data && d_idx && d_w && d_out, "s3_interp_gather: NULL argument");
    S3_REQUIRE(k >= 1 && k <= 64, "s3_interp_gather: k=%d out of range", k);
    S3_REQUIRE(n_src >= 1 && row_len >= 0 && n_cells >= 0, "s3_interp_gather: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (data_dtype == S3_F32 && out_dtype == S3_F32)
        return launch_interp<float, float, float, 0>(d_data, row_len, d_idx, d_w, n_cells, k, d_out_row, d_out, st);
    if (data_dtype == S3_F32 && out_dtype == S3_F64)
        return launch_interp<float, double, double, 1>(d_data, row_len, d_idx, d_w, n_cells, k, d_out_row, d_out, st);
    if (data_dtype == S3_F64 && out_dtype == S3_F64)
        return launch_interp<double, double, double, 1>(d_data, row_len, d_idx, d_w, n_cells, k, d_out_row, d_out, st);
    s3::set_error("s3_interp_gather: unsupported dtype combination data=%d out=%d", data_dtype, out_dtype);
    return S3_ERR_UNSUPPORTED;
}
