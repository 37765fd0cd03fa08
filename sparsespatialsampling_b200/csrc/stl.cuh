// Point-in-mesh test of GeometrySTL3D for whole batches of cell nodes (geometry_STL_3d.py:81-124,
// pyvista select_enclosed_points(check_surface=False, tolerance=0.001)).
//
// One thread per NODE (not per cell), 256 nodes per CTA. The triangle list of the geometry is stored in Morton order of
// the triangle centroids and cut into tiles of kStlTile triangles with one bounding box per tile (host side,
// geometry/surfaces.py). A CTA
//   1. computes the bounding box of its own nodes (block reduction),
//   2. lists the tiles that can matter for ANY of its nodes: the +x ray of a node can only cross triangles whose y-z
//      extent contains the node and whose x extent reaches beyond it, the tolerance band only triangles whose box
//      (grown by the tolerance) contains it,
//   3. streams those tiles through shared memory (cp.async, two stages), computes the per-triangle boxes once per tile,
//      and every thread tests its node against the staged triangles: the tile's box, then the triangle's box, then
//      exactly the arithmetic of
//      the single-point test in geometry.cuh (in_stl: closest-point distance for the tolerance band, 2-D edge functions
//      and the x of the hit for the ray) -- the result is the same predicate, order independent (a crossing COUNT and
//      an OR), hence bit-identical to in_stl and to the CPU oracle.
// Parameter block of the geometry: lo[3], hi[3], tol, n_tri * 9 doubles (Morton order), n_tiles * 6 doubles (tile boxes).
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace s3 {

constexpr int kStlTile = 128;
constexpr int kStlThreads = 256;

struct StlPoints {
    // mode 0: nodes of cells (centre +- half width), mode 1: explicit points [n_pts, 3]
    int mode;
    const double* center;      // mode 0: [cap, 3]
    const int32_t* level;      // mode 0
    const int64_t* cells;      // mode 0: optional cell list
    int64_t first;             // mode 0
    double width;              // mode 0
    const double* points;      // mode 1
};

__device__ __forceinline__ void stl_dir(int c, int* d) {          // CH order (s_cube.py:29, :188-194), 3-D
    const int dx[4] = {-1, -1, 1, 1};
    const int dy[4] = {-1, 1, 1, -1};
    d[0] = dx[c & 3];
    d[1] = dy[c & 3];
    d[2] = (c < 4) ? 1 : -1;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(kStlThreads)
stl_inside_kernel(const StlPoints src, int64_t n_pts, const double* __restrict__ par, int n_tri,
                  uint8_t* __restrict__ inside) {
    __shared__ double s_tri[2][kStlTile * 9];
    __shared__ double s_tbox[kStlTile][6];
    __shared__ double s_red[6][kStlThreads / 32];
    __shared__ double s_box[6];
    __shared__ int s_tiles[1024];
    __shared__ int s_ntiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = (int64_t)blockIdx.x * kStlThreads + tid;
    const double* lo = par;
    const double* hi = par + 3;
    const double tol = par[6];
    const double* tri = par + 7;
    const int n_tiles = (n_tri + kStlTile - 1) / kStlTile;
    const double* tile_box = tri + (size_t)n_tri * 9;

    // ---- this thread's node
    double p[3] = {0.0, 0.0, 0.0};
    bool active = q < n_pts;
    if (active) {
        if (src.mode == 0) {
            const int64_t t = q >> 3;
            const int j = (int)(q & 7);
            const int64_t cell = src.cells ? src.cells[t] : src.first + t;
            const double h = ldexp(src.width, -(src.level[cell] + 1));
            int d[3];
            stl_dir(j, d);
            for (int a = 0; a < 3; ++a) p[a] = __dadd_rn(src.center[cell * 3 + a], d[a] > 0 ? h : -h);
        } else {
            for (int a = 0; a < 3; ++a) p[a] = src.points[q * 3 + a];
        }
        // in_stl's first test: outside the bounding box grown by the tolerance -> outside
        for (int a = 0; a < 3; ++a)
            if (p[a] < lo[a] - tol || p[a] > hi[a] + tol) active = false;
    }
    // ---- bounding box of the CTA's candidate nodes
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        bmin[a] = active ? p[a] : 1e300;
        bmax[a] = active ? p[a] : -1e300;
        for (int o = 16; o > 0; o >>= 1) {
            bmin[a] = fmin(bmin[a], __shfl_xor_sync(0xffffffffu, bmin[a], o));
            bmax[a] = fmax(bmax[a], __shfl_xor_sync(0xffffffffu, bmax[a], o));
        }
        if (lane == 0) { s_red[a][warp] = bmin[a]; s_red[3 + a][warp] = bmax[a]; }
    }
    if (tid == 0) s_ntiles = 0;
    __syncthreads();
    if (tid < 6) {
        double v = s_red[tid][0];
        for (int w = 1; w < kStlThreads / 32; ++w) v = tid < 3 ? fmin(v, s_red[tid][w]) : fmax(v, s_red[tid][w]);
        s_box[tid] = v;
    }
    __syncthreads();
    const bool any = s_box[0] <= s_box[3];
    if (!any) {                                          // no node of this CTA is near the geometry
        if (q < n_pts) inside[q] = 0;
        return;
    }
    // ---- tiles that can matter for some node of the CTA (order is irrelevant: the result is a count and an OR)
    const double ext = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const double jit = ext * 2e-9;                       // the ray's y/z perturbation (< 1.8e-9 * ext), conservatively
    for (int t0 = 0; t0 < n_tiles; t0 += (int)(sizeof(s_tiles) / sizeof(int))) {
        __syncthreads();
        if (tid == 0) s_ntiles = 0;
        __syncthreads();
        const int t_end = min(n_tiles, t0 + (int)(sizeof(s_tiles) / sizeof(int)));
        for (int t = t0 + tid; t < t_end; t += kStlThreads) {
            const double* b = tile_box + (size_t)t * 6;
            const bool yz = b[1] - tol - jit <= s_box[4] && b[4] + tol + jit >= s_box[1] &&
                            b[2] - tol - jit <= s_box[5] && b[5] + tol + jit >= s_box[2];
            const bool x_reach = b[3] + tol >= s_box[0];
            if (yz && x_reach) s_tiles[atomicAdd(&s_ntiles, 1)] = t;
        }
        __syncthreads();
        const int n_list = s_ntiles;
        // ---- stream the listed tiles: cp.async into stage (i & 1) while stage ((i - 1) & 1) is being tested
        auto stage_tile = [&](int i) {
            const int t = s_tiles[i];
            const int nt = min(kStlTile, n_tri - t * kStlTile);
            const double* g = tri + (size_t)t * kStlTile * 9;
            for (int e = tid; e < nt * 9; e += kStlThreads) cp_async8(&s_tri[i & 1][e], g + e);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (n_list > 0) stage_tile(0);
        const double py = p[1] + ext * 1.4142135623730951e-9;
        const double pz = p[2] + ext * 1.7320508075688772e-9;
        int crossings = 0;
        bool near = false;
        // (crossings / near accumulate across the outer t0 loop through registers declared below)
        for (int i = 0; i < n_list; ++i) {
            if (i + 1 < n_list) {
                stage_tile(i + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            const int t = s_tiles[i];
            const int nt = min(kStlTile, n_tri - t * kStlTile);
            const double* st = s_tri[i & 1];
            if (tid < nt) {                              // per-triangle boxes, once per tile
                const double* a = st + 9 * tid;
                for (int c = 0; c < 3; ++c) {
                    s_tbox[tid][c] = fmin(fmin(a[c], a[3 + c]), a[6 + c]);
                    s_tbox[tid][3 + c] = fmax(fmax(a[c], a[3 + c]), a[6 + c]);
                }
            }
            __syncthreads();
            // this node against the tile's own box first: most (node, tile) pairs of a CTA end here
            const double* gb = tile_box + (size_t)t * 6;
            const bool tile_yz = py >= gb[1] && py <= gb[4] && pz >= gb[2] && pz <= gb[5] && gb[3] > p[0];
            const bool tile_band = p[0] >= gb[0] - tol && p[0] <= gb[3] + tol && p[1] >= gb[1] - tol &&
                                   p[1] <= gb[4] + tol && p[2] >= gb[2] - tol && p[2] <= gb[5] + tol;
            if (active && (tile_yz || (tile_band && !near))) {
                for (int k = 0; k < nt; ++k) {
                    const double* tb = s_tbox[k];
                    const bool in_yz = py >= tb[1] && py <= tb[4] && pz >= tb[2] && pz <= tb[5];
                    const bool band = p[0] >= tb[0] - tol && p[0] <= tb[3] + tol && p[1] >= tb[1] - tol &&
                                      p[1] <= tb[4] + tol && p[2] >= tb[2] - tol && p[2] <= tb[5] + tol;
                    if (!in_yz && !band) continue;
                    const double* a = st + 9 * k;
                    const double* b = a + 3;
                    const double* c = a + 6;
                    if (band && !near && pt_tri_dist2(p, a, b, c) <= tol * tol) near = true;
                    if (in_yz && tb[3] > p[0] && stl_ray_crosses(p[0], py, pz, a, b, c)) ++crossings;
                }
            }
            __syncthreads();                             // stage (i & 1) and s_tbox may be overwritten now
        }
        if (q < n_pts) {
            // partial result of this batch of tiles: combine across batches through global memory (n_tiles <= 1024:
            // a single batch, the common case -- 131 072 triangles)
            const uint8_t prev = t0 > 0 ? inside[q] : 0;
            const uint8_t now = (uint8_t)((near ? 2 : 0) | (crossings & 1));
            inside[q] = (uint8_t)((prev | (now & 2)) ^ (now & 1));
        }
    }
    __syncthreads();
    if (q < n_pts) {
        const uint8_t v = active ? inside[q] : 0;
        inside[q] = (uint8_t)(((v & 2) || (v & 1)) ? 1 : 0);
    }
}

}  // namespace s3
