// KNN index build + query entry points (see knn.cuh for the algorithm and the parity notes).
#include "knn.cuh"
#include "radix_sort.cuh"
#include "../../include/s3b200.h"

namespace s3 {

// ------------------------------------------------------------------ build kernels
__global__ void __launch_bounds__(256)
bbox_partial_kernel(const double* __restrict__ coords, int64_t n, int dim, double* __restrict__ partial) {
    // partial layout: [gridDim.x][2*dim] = (min_0..min_{d-1}, max_0..max_{d-1})
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        for (int a = 0; a < dim; ++a) {
            double v = coords[i * dim + a];
            mn[a] = fmin(mn[a], v);
            mx[a] = fmax(mx[a], v);
        }
    }
    __shared__ double smn[3][8], smx[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = 0; a < dim; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], shfl_xor_d(mn[a], o));
            mx[a] = fmax(mx[a], shfl_xor_d(mx[a], o));
        }
        if (lane == 0) { smn[a][warp] = mn[a]; smx[a][warp] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int a = 0; a < dim; ++a) {
            double m0 = smn[a][0], m1 = smx[a][0];
            for (int w = 1; w < 8; ++w) { m0 = fmin(m0, smn[a][w]); m1 = fmax(m1, smx[a][w]); }
            partial[(int64_t)blockIdx.x * 2 * dim + a] = m0;
            partial[(int64_t)blockIdx.x * 2 * dim + dim + a] = m1;
        }
    }
}

__global__ void bbox_final_kernel(const double* __restrict__ partial, int nparts, int dim, double* __restrict__ bb) {
    // single warp
    const int lane = threadIdx.x;
    for (int a = 0; a < dim; ++a) {
        double mn = 1e300, mx = -1e300;
        for (int p = lane; p < nparts; p += 32) {
            mn = fmin(mn, partial[(int64_t)p * 2 * dim + a]);
            mx = fmax(mx, partial[(int64_t)p * 2 * dim + dim + a]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, shfl_xor_d(mn, o));
            mx = fmax(mx, shfl_xor_d(mx, o));
        }
        if (lane == 0) { bb[a] = mn; bb[dim + a] = mx; }
    }
}

__device__ __forceinline__ uint64_t spread2(uint64_t v) {  // 32 bits -> every 2nd bit
    v &= 0xffffffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
__device__ __forceinline__ uint64_t spread3(uint64_t v) {  // 21 bits -> every 3rd bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// Space-filling-curve key of a point. g_curve = 1 (default): Hilbert curve (Skilling's transpose algorithm) -- every run
// of consecutive points on it is one compact blob, so the 32-point leaves and the 32-leaf nodes of the implicit
// hierarchy have tight boxes. On the Z (Morton) curve (g_curve = 0) a run that crosses an octant boundary of a high
// level consists of two far-apart clusters: its box spans the gap, intersects almost every query ball nearby and is
// never pruned -- the k = 26 search on 10 M points popped ~800 nodes per query that way (profiles/r1_gain_kernel_c4).
int g_curve = 1;

template <int DIM>
__device__ __forceinline__ void hilbert_transpose(uint64_t* x, int bits) {
    const uint64_t m = 1ull << (bits - 1);
    for (uint64_t q = m; q > 1; q >>= 1) {
        const uint64_t p = q - 1;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            if (x[i] & q) x[0] ^= p;
            else { const uint64_t t = (x[0] ^ x[i]) & p; x[0] ^= t; x[i] ^= t; }
        }
    }
#pragma unroll
    for (int i = 1; i < DIM; ++i) x[i] ^= x[i - 1];
    uint64_t t = 0;
    for (uint64_t q = m; q > 1; q >>= 1)
        if (x[DIM - 1] & q) t ^= q - 1;
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] ^= t;
}

__global__ void __launch_bounds__(256)
morton_kernel(const double* __restrict__ coords, int64_t n, int dim, const double* __restrict__ bb,
              uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int curve) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bits = dim == 2 ? 31 : 21;
    const double scale = (double)(1ull << bits);
    uint64_t q[3] = {0, 0, 0};
    for (int a = 0; a < dim; ++a) {
        const double lo = bb[a], ext = bb[dim + a] - bb[a];
        double u = ext > 0.0 ? (coords[i * dim + a] - lo) / ext : 0.0;
        double f = u * scale;
        uint64_t c = f <= 0.0 ? 0ull : (uint64_t)f;
        if (c > (1ull << bits) - 1) c = (1ull << bits) - 1;
        q[a] = c;
    }
    uint64_t key;
    if (curve == 1) {
        if (dim == 2) { hilbert_transpose<2>(q, bits); key = (spread2(q[0]) << 1) | spread2(q[1]); }
        else { hilbert_transpose<3>(q, bits); key = (spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]); }
    } else {
        key = dim == 2 ? (spread2(q[0]) | (spread2(q[1]) << 1))
                       : (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2));
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
gather_sorted_kernel(const double* __restrict__ coords, const uint32_t* __restrict__ perm, int64_t n, int64_t n_pad,
                     int dim, double* __restrict__ pts, int32_t* __restrict__ orig) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i < n) {
        const uint32_t p = perm[i];
        for (int a = 0; a < dim; ++a) pts[(int64_t)a * n_pad + i] = coords[(int64_t)p * dim + a];
        orig[i] = (int32_t)p;
    } else {
        for (int a = 0; a < dim; ++a) pts[(int64_t)a * n_pad + i] = 0.0;
        orig[i] = 0x7fffffff;
    }
}

// One warp per node of `level`; children are points (level 0) or nodes of level-1.
__global__ void __launch_bounds__(256)
build_boxes_kernel(KnnView ix, int level, double* __restrict__ box_lo, double* __restrict__ box_hi) {
    const int64_t node = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= ix.level_count[level]) return;
    for (int a = 0; a < ix.dim; ++a) {
        double mn = 1e300, mx = -1e300;
        const int64_t c = node * kKnnFan + lane;
        if (level == 0) {
            if (c < ix.n) { mn = mx = ix.pts[(int64_t)a * ix.n_pad + c]; }
        } else if (c < ix.level_count[level - 1]) {
            const int64_t o = ix.level_offset[level - 1] + c;
            mn = box_lo[(int64_t)a * ix.n_nodes + o];
            mx = box_hi[(int64_t)a * ix.n_nodes + o];
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, shfl_xor_d(mn, o));
            mx = fmax(mx, shfl_xor_d(mx, o));
        }
        if (lane == 0) {
            box_lo[(int64_t)a * ix.n_nodes + ix.level_offset[level] + node] = mn;
            box_hi[(int64_t)a * ix.n_nodes + ix.level_offset[level] + node] = mx;
        }
    }
}

// ------------------------------------------------------------------ query kernels
constexpr int kQueryWarps = 4;

template <int DIM, int MODE>
// MODE 0: kneighbors (idx int64 + dist fp64); 1: IDW predict; 2: export tables (idx int32, w fp32, w fp64)
__global__ void __launch_bounds__(kQueryWarps * 32)
knn_query_kernel(KnnView ix, const double* __restrict__ query, int64_t nq, int k, int64_t* __restrict__ out_idx,
                 double* __restrict__ out_dist, double* __restrict__ out_pred, int32_t* __restrict__ tab_idx,
                 float* __restrict__ tab_w32, double* __restrict__ tab_w64) {
    __shared__ double s_lb[kQueryWarps][kKnnStack];
    __shared__ int32_t s_node[kQueryWarps][kKnnStack];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * kQueryWarps + warp;
    if (qi >= nq) return;
    double q[3];
#pragma unroll
    for (int a = 0; a < DIM; ++a) q[a] = query[qi * DIM + a];
    WarpTopK best = warp_knn_search<DIM>(ix, q, k, s_lb[warp], s_node[warp]);
    if (MODE == 0) {
        if (lane < k) {
            out_idx[qi * k + lane] = (int64_t)best.i;
            out_dist[qi * k + lane] = __dsqrt_rn(best.d);
        }
    } else if (MODE == 1) {
        const double pred = warp_idw_predict(best, k, ix.values);
        if (lane == 0) out_pred[qi] = pred;
    } else {
        // reference: sparseSpatialSampling/export.py:428-429  w = 1/clamp(dist, 1e-12); w /= w.sum(1)
        const bool act = lane < k;
        double dist = act ? __dsqrt_rn(best.d) : 1.0;
        dist = dist < 1e-12 ? 1e-12 : dist;
        double w = act ? __ddiv_rn(1.0, dist) : 0.0;
        double sum;
        if (k == 8) {
            // torch fp64 sum(dim=1) over 8 contiguous columns: (((a0+a4)+(a1+a5))+(a2+a6))+(a3+a7)
            double v = __dadd_rn(w, shfl_d(w, (lane & 3) + 4));  // lanes 0..3 hold a_j + a_{j+4}
            double s01 = __dadd_rn(shfl_d(v, 0), shfl_d(v, 1));
            double s012 = __dadd_rn(s01, shfl_d(v, 2));
            sum = __dadd_rn(s012, shfl_d(v, 3));
        } else {
            sum = shfl_d(warp_numpy_pairwise_sum(w, k), 0);
        }
        const double wn = __ddiv_rn(w, sum);
        if (act) {
            tab_idx[qi * k + lane] = best.i;
            tab_w32[qi * k + lane] = (float)wn;
            if (tab_w64) tab_w64[qi * k + lane] = wn;
        }
    }
}

template <int MODE>
static int launch_query(const KnnIndex& ix, const double* query, int64_t nq, int k, int64_t* out_idx,
                        double* out_dist, double* out_pred, int32_t* tab_idx, float* tab_w32, double* tab_w64,
                        cudaStream_t stream) {
    S3_REQUIRE(k >= 1 && k <= kKnnMaxK, "k=%d out of range [1,%d]", k, kKnnMaxK);
    S3_REQUIRE(k <= ix.n, "k=%d larger than the number of indexed points %lld", k, (long long)ix.n);
    if (nq == 0) return S3_OK;
    const int64_t blocks = ceil_div(nq, kQueryWarps);
    S3_REQUIRE(blocks < ((int64_t)1 << 31), "too many queries");
    KnnView v = make_view(ix);
    if (ix.dim == 2)
        knn_query_kernel<2, MODE><<<(unsigned)blocks, kQueryWarps * 32, 0, stream>>>(v, query, nq, k, out_idx, out_dist,
                                                                                    out_pred, tab_idx, tab_w32, tab_w64);
    else
        knn_query_kernel<3, MODE><<<(unsigned)blocks, kQueryWarps * 32, 0, stream>>>(v, query, nq, k, out_idx, out_dist,
                                                                                    out_pred, tab_idx, tab_w32, tab_w64);
    S3_LAUNCH_CHECK();
    note_launch(1);
    return S3_OK;
}

static int knn_build_impl(const double* coords, int64_t n, int dim, const double* values, cudaStream_t stream,
                          KnnIndex* ix) {
    ix->dim = dim;
    ix->n = n;
    ix->n_pad = ceil_div(n, kKnnFan) * kKnnFan;
    // level sizes (the search encodes a node index in 24 bits of a stack entry: at most 2^24 leaves = 5.4e8 points)
    S3_REQUIRE(ix->n_pad / kKnnFan < ((int64_t)1 << 24), "s3_knn_build: more than 2^24 leaves (%lld points)", (long long)n);
    int64_t cnt = ix->n_pad / kKnnFan, off = 0;
    int L = 0;
    while (true) {
        S3_REQUIRE(L < kKnnMaxLevels, "too many tree levels");
        ix->level_count[L] = cnt;
        ix->level_offset[L] = off;
        off += cnt;
        ++L;
        if (cnt == 1) break;
        cnt = ceil_div(cnt, kKnnFan);
    }
    for (int i = L; i < kKnnMaxLevels; ++i) { ix->level_count[i] = 0; ix->level_offset[i] = 0; }
    ix->n_levels = L;
    ix->n_nodes = off;

    S3_CUDA(cudaMalloc(&ix->pts, sizeof(double) * dim * ix->n_pad));
    S3_CUDA(cudaMalloc(&ix->orig, sizeof(int32_t) * ix->n_pad));
    S3_CUDA(cudaMalloc(&ix->box_lo, sizeof(double) * dim * ix->n_nodes));
    S3_CUDA(cudaMalloc(&ix->box_hi, sizeof(double) * dim * ix->n_nodes));
    if (values) {
        S3_CUDA(cudaMalloc(&ix->values, sizeof(double) * n));
        S3_CUDA(cudaMemcpyAsync(ix->values, values, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    }

    Scratch scratch(stream);
    const int nparts = (int)(ceil_div(n, 256) < 1024 ? ceil_div(n, 256) : 1024);
    double *partial = nullptr, *bb = nullptr;
    uint64_t *ka = nullptr, *kb = nullptr;
    uint32_t *va = nullptr, *vb = nullptr;
    S3_TRY(scratch.alloc(&partial, (size_t)nparts * 2 * dim));
    S3_TRY(scratch.alloc(&bb, 2 * dim));
    S3_TRY(scratch.alloc(&ka, n));
    S3_TRY(scratch.alloc(&kb, n));
    S3_TRY(scratch.alloc(&va, n));
    S3_TRY(scratch.alloc(&vb, n));

    bbox_partial_kernel<<<nparts, 256, 0, stream>>>(coords, n, dim, partial);
    bbox_final_kernel<<<1, 32, 0, stream>>>(partial, nparts, dim, bb);
    morton_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(coords, n, dim, bb, ka, va, g_curve);
    S3_LAUNCH_CHECK();
    bool in_a = true;
    S3_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, 64, stream, &in_a));
    const uint32_t* perm = in_a ? va : vb;
    gather_sorted_kernel<<<(unsigned)ceil_div(ix->n_pad, 256), 256, 0, stream>>>(coords, perm, n, ix->n_pad, dim, ix->pts,
                                                                               ix->orig);
    S3_LAUNCH_CHECK();
    KnnView v = make_view(*ix);
    for (int l = 0; l < L; ++l) {
        const int64_t threads = ix->level_count[l] * 32;
        build_boxes_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, stream>>>(v, l, ix->box_lo, ix->box_hi);
    }
    S3_LAUNCH_CHECK();
    double hbb[6];
    S3_CUDA(cudaMemcpyAsync(hbb, bb, sizeof(double) * 2 * dim, cudaMemcpyDeviceToHost, stream));
    S3_CUDA(cudaStreamSynchronize(stream));
    for (int a = 0; a < dim; ++a) { ix->bb_lo[a] = hbb[a]; ix->bb_hi[a] = hbb[dim + a]; }
    note_launch(28 + L);
    return S3_OK;
}

static void knn_release(KnnIndex* ix) {
    if (!ix) return;
    cudaFree(ix->pts);
    cudaFree(ix->orig);
    cudaFree(ix->box_lo);
    cudaFree(ix->box_hi);
    cudaFree(ix->values);
    delete ix;
}

}  // namespace s3

using namespace s3;

struct s3_knn {
    KnnIndex ix;
};

extern "C" {

int s3_knn_build(const double* d_coords, int64_t n, int dim, const double* d_values, void* stream, s3_knn_t** out) {
    S3_REQUIRE(out != nullptr, "s3_knn_build: out is NULL");
    *out = nullptr;
    S3_REQUIRE(d_coords != nullptr, "s3_knn_build: coords is NULL");
    S3_REQUIRE(dim == 2 || dim == 3, "s3_knn_build: dim must be 2 or 3, got %d", dim);
    S3_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), "s3_knn_build: n=%lld out of range", (long long)n);
    s3_knn* h = new s3_knn();
    int rc = knn_build_impl(d_coords, n, dim, d_values, (cudaStream_t)stream, &h->ix);
    if (rc != S3_OK) {
        cudaFree(h->ix.pts); cudaFree(h->ix.orig); cudaFree(h->ix.box_lo); cudaFree(h->ix.box_hi); cudaFree(h->ix.values);
        delete h;
        return rc;
    }
    *out = h;
    return S3_OK;
}

int s3_knn_free(s3_knn_t* h) {
    if (!h) return S3_OK;
    cudaFree(h->ix.pts);
    cudaFree(h->ix.orig);
    cudaFree(h->ix.box_lo);
    cudaFree(h->ix.box_hi);
    cudaFree(h->ix.values);
    delete h;
    return S3_OK;
}

int s3_knn_query(const s3_knn_t* h, const double* d_query, int64_t nq, int k, int64_t* d_idx, double* d_dist,
                 void* stream) {
    S3_REQUIRE(h && nq >= 0, "s3_knn_query: NULL handle or negative query count");
    if (nq == 0) return S3_OK;                               // no queries: the (empty) buffers may be NULL
    S3_REQUIRE(d_query && d_idx && d_dist, "s3_knn_query: NULL argument");
    return launch_query<0>(h->ix, d_query, nq, k, d_idx, d_dist, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int s3_knn_predict(const s3_knn_t* h, const double* d_query, int64_t nq, int k, double* d_pred, void* stream) {
    S3_REQUIRE(h && nq >= 0, "s3_knn_predict: NULL handle or negative query count");
    if (nq == 0) return S3_OK;
    S3_REQUIRE(d_query && d_pred, "s3_knn_predict: NULL argument");
    S3_REQUIRE(h->ix.values != nullptr, "s3_knn_predict: index was built without values");
    return launch_query<1>(h->ix, d_query, nq, k, nullptr, nullptr, d_pred, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int s3_knn_tables(const s3_knn_t* h, const double* d_query, int64_t nq, int k, int32_t* d_idx, float* d_w32,
                  double* d_w64, void* stream) {
    S3_REQUIRE(h && nq >= 0, "s3_knn_tables: NULL handle or negative query count");
    if (nq == 0) return S3_OK;
    S3_REQUIRE(d_query && d_idx && d_w32, "s3_knn_tables: NULL argument");
    return launch_query<2>(h->ix, d_query, nq, k, nullptr, nullptr, nullptr, d_idx, d_w32, d_w64, (cudaStream_t)stream);
}

int s3_morton_order(const double* d_coords, int64_t n, int dim, int32_t* d_perm, void* stream) {
    S3_REQUIRE(dim == 2 || dim == 3, "s3_morton_order: dim must be 2 or 3");
    S3_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "s3_morton_order: n out of range");
    if (n == 0) return S3_OK;
    S3_REQUIRE(d_coords && d_perm, "s3_morton_order: NULL argument");
    cudaStream_t st_ = (cudaStream_t)stream;
    Scratch scratch(st_);
    const int nparts = (int)(ceil_div(n, 256) < 1024 ? ceil_div(n, 256) : 1024);
    double *partial = nullptr, *bb = nullptr;
    uint64_t *ka = nullptr, *kb = nullptr;
    uint32_t *va = nullptr, *vb = nullptr;
    S3_TRY(scratch.alloc(&partial, (size_t)nparts * 2 * dim));
    S3_TRY(scratch.alloc(&bb, 2 * dim));
    S3_TRY(scratch.alloc(&ka, n));
    S3_TRY(scratch.alloc(&kb, n));
    S3_TRY(scratch.alloc(&va, n));
    S3_TRY(scratch.alloc(&vb, n));
    bbox_partial_kernel<<<nparts, 256, 0, st_>>>(d_coords, n, dim, partial);
    bbox_final_kernel<<<1, 32, 0, st_>>>(partial, nparts, dim, bb);
    morton_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st_>>>(d_coords, n, dim, bb, ka, va, g_curve);
    S3_LAUNCH_CHECK();
    bool in_a = true;
    S3_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, 64, st_, &in_a));
    S3_CUDA(cudaMemcpyAsync(d_perm, in_a ? va : vb, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st_));
    note_launch(27);
    return S3_OK;
}

}  // extern "C"


namespace s3 {
const KnnIndex* knn_index_of(const s3_knn_t* h) { return h ? &h->ix : nullptr; }
}
