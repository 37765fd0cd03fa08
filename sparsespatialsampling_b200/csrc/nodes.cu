// Final grid assembly: unique vertex table + per-cell node ids of the leaf cells.
//
// Stands in for the node bookkeeping of the reference (_assign_indices s_cube.py:1188-1536,
// _resort_nodes_and_indices_of_grid :734-772, numba renumber_node_indices_parallel :1695-1736).
// The reference shares a node only between same-level leaf neighbours that exist when the node is
// created, so its vertex table depends on the refinement history (and may hold duplicate coordinates
// at level transitions). Here every corner is identified by its integer position on the finest
// lattice, so corners are shared whenever they coincide. Per-cell corner coordinates agree with the
// reference to rounding (a shared corner keeps the coordinates computed from the first leaf using it).
#include "common.cuh"
#include "radix_sort.cuh"
#include "../../include/s3b200.h"

namespace s3 {

__device__ __forceinline__ void node_direction(int dim, int c, int* d) {
    const int dx[4] = {-1, -1, 1, 1};
    const int dy[4] = {-1, 1, 1, -1};
    d[0] = dx[c & 3];
    d[1] = dy[c & 3];
    if (dim == 3) d[2] = (c < 4) ? 1 : -1;
}

__global__ void __launch_bounds__(256)
node_keys_kernel(const int64_t* __restrict__ leaves, int64_t n_leaves, const int32_t* __restrict__ level,
                 const int32_t* __restrict__ lattice, int dim, int max_level, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ vals) {
    const int nn = 1 << dim;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_leaves * nn) return;
    const int64_t cell = leaves[t / nn];
    const int j = (int)(t % nn);
    int d[3];
    node_direction(dim, j, d);
    const int sh = max_level - level[cell];
    uint64_t key = 0;
    const int bits = dim == 2 ? 32 : 21;
    for (int a = 0; a < dim; ++a) {
        const uint64_t c = ((uint64_t)(lattice[cell * dim + a] + (d[a] > 0 ? 1 : 0))) << sh;
        key |= c << (a * bits);
    }
    keys[t] = key;
    vals[t] = (uint32_t)t;
}

__global__ void __launch_bounds__(256)
mark_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ heads) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    heads[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// three-phase exclusive scan (tile sums -> scan of tile sums -> tile rescan)
constexpr int kScanTile = 2048;

__global__ void __launch_bounds__(256)
scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sw[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t acc = 0;
    for (int j = threadIdx.x; j < kScanTile; j += 256)
        if (base + j < n) acc += in[base + j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < 8; ++w) s += sw[w];
        sums[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
scan_tile_apply_kernel(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ tile_off,
                       uint32_t* __restrict__ out) {
    // each thread owns 8 consecutive items of the tile
    __shared__ uint32_t sw[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        v[e] = (base + e < n) ? in[base + e] : 0u;
        s += v[e];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += sw[w];
    uint32_t run = tile_off[blockIdx.x] + woff + inc - s;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (base + e < n) out[base + e] = run;
        run += v[e];
    }
}

static int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t stream, Scratch& scratch) {
    const int64_t tiles = ceil_div(n, kScanTile);
    uint32_t* sums = nullptr;
    S3_TRY(scratch.alloc(&sums, tiles));
    scan_tile_sums_kernel<<<(unsigned)tiles, 256, 0, stream>>>(in, n, sums);
    if (tiles <= 1024 * 1024) {
        radix_scan_kernel<<<1, 1024, 0, stream>>>(sums, tiles);
    } else {
        set_error("exclusive_scan_u32: input too large");
        return S3_ERR_UNSUPPORTED;
    }
    scan_tile_apply_kernel<<<(unsigned)tiles, 256, 0, stream>>>(in, n, sums, out);
    S3_LAUNCH_CHECK();
    note_launch(3);
    return S3_OK;
}

__global__ void __launch_bounds__(256)
node_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ heads,
                 const uint32_t* __restrict__ uid, int64_t n, const int64_t* __restrict__ leaves,
                 const double* __restrict__ center, const int32_t* __restrict__ level, int dim, double width,
                 int32_t* __restrict__ faces, double* __restrict__ vertices, uint32_t* __restrict__ n_unique) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = vals[i];
    const uint32_t id = uid[i];
    faces[v] = (int32_t)id;
    if (heads[i]) {
        const int nn = 1 << dim;
        const int64_t cell = leaves[v / nn];
        const int j = (int)(v % nn);
        int d[3];
        node_direction(dim, j, d);
        const double h = ldexp(width, -(level[cell] + 1));
        for (int a = 0; a < dim; ++a) vertices[(int64_t)id * dim + a] = __dadd_rn(center[cell * dim + a], d[a] > 0 ? h : -h);
    }
    if (i == n - 1) *n_unique = id + 1;
}

__global__ void node_fix_uid_kernel(const uint32_t* __restrict__ heads, uint32_t* __restrict__ uid, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !heads[i]) uid[i] -= 1u;
}

}  // namespace s3

using namespace s3;

extern "C" int s3_build_nodes(const int64_t* d_leaves, int64_t n_leaves, const double* d_center,
                              const int32_t* d_level, const int32_t* d_lattice, int dim, int max_level,
                              double width, int32_t* d_faces, double* d_vertices, int64_t* n_vertices,
                              void* stream) {
    S3_REQUIRE(d_leaves && d_center && d_level && d_lattice && d_faces && d_vertices && n_vertices,
               "s3_build_nodes: NULL argument");
    S3_REQUIRE(dim == 2 || dim == 3, "s3_build_nodes: dim must be 2 or 3");
    S3_REQUIRE(max_level + 1 <= (dim == 2 ? 31 : 20), "s3_build_nodes: refinement level %d too deep for the node keys",
               max_level);
    *n_vertices = 0;
    if (n_leaves == 0) return S3_OK;
    cudaStream_t st_ = (cudaStream_t)stream;
    const int64_t n = n_leaves << dim;
    S3_REQUIRE(n < ((int64_t)1 << 31), "s3_build_nodes: too many cells");
    Scratch scratch(st_);
    uint64_t *ka = nullptr, *kb = nullptr;
    uint32_t *va = nullptr, *vb = nullptr, *heads = nullptr, *uid = nullptr, *d_nu = nullptr;
    S3_TRY(scratch.alloc(&ka, n));
    S3_TRY(scratch.alloc(&kb, n));
    S3_TRY(scratch.alloc(&va, n));
    S3_TRY(scratch.alloc(&vb, n));
    S3_TRY(scratch.alloc(&heads, n));
    S3_TRY(scratch.alloc(&uid, n));
    S3_TRY(scratch.alloc(&d_nu, 1));
    node_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st_>>>(d_leaves, n_leaves, d_level, d_lattice, dim, max_level,
                                                                 ka, va);
    S3_LAUNCH_CHECK();
    bool in_a = true;
    const int key_bits = 64;
    S3_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, key_bits, st_, &in_a));
    const uint64_t* ks = in_a ? ka : kb;
    const uint32_t* vs = in_a ? va : vb;
    mark_heads_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st_>>>(ks, n, heads);
    S3_TRY(exclusive_scan_u32(heads, uid, n, st_, scratch));
    // the exclusive scan counts the heads strictly before i: a head's own id is that count, a non-head
    // belongs to the previous head (count - 1)
    node_fix_uid_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st_>>>(heads, uid, n);
    node_emit_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st_>>>(ks, vs, heads, uid, n, d_leaves, d_center, d_level, dim,
                                                                 width, d_faces, d_vertices, d_nu);
    S3_LAUNCH_CHECK();
    note_launch(30);
    uint32_t nu = 0;
    S3_CUDA(cudaMemcpyAsync(&nu, d_nu, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
    S3_CUDA(cudaStreamSynchronize(st_));
    *n_vertices = (int64_t)nu;
    return S3_OK;
}

