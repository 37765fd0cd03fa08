// Exact k-nearest-neighbour search over the original CFD point cloud.
//
// Replaces sklearn's KD-tree behind KNeighborsRegressor / NearestNeighbors as used by the reference
// (sparseSpatialSampling/s_cube.py:161-163,224,328,372; sparseSpatialSampling/export.py:120,423-441).
//
// Index: points sorted along a Morton curve, buckets of 32 consecutive points form the leaves of an
// implicit 32-ary bounding-box hierarchy (one warp lane per child / per point).
// Query: one warp per query point, best-first descent with a shared-memory stack. Distances are the
// reduced (squared) Euclidean distance accumulated dimension by dimension WITHOUT fused multiply-add,
// i.e. exactly the value sklearn's tree ranks by; a subtree is pruned only if its box lower bound
// (computed with the same monotone operations) is strictly greater than the current k-th distance, so
// the result is the exact top-k of the computed fp64 values. Ties are resolved by the smaller original
// point index (sklearn leaves ties unspecified).
#pragma once
#include "common.cuh"

namespace s3 {

constexpr int kKnnFan = 32;        // children per node == points per leaf == warp size
constexpr int kKnnMaxLevels = 8;   // 32^8 points
constexpr int kKnnMaxK = 32;
constexpr int kKnnStack = 32 * kKnnMaxLevels;

struct KnnIndex {
    int dim = 0;
    int64_t n = 0;
    int n_levels = 0;                      // levels of boxes; level 0 = leaves
    int64_t level_count[kKnnMaxLevels];    // nodes per level
    int64_t level_offset[kKnnMaxLevels];   // offset of the level inside the box arrays
    int64_t n_nodes = 0;
    double* pts = nullptr;                 // [dim][n_pad] sorted coordinates (SoA)
    int32_t* orig = nullptr;               // [n_pad] original index of the sorted point
    double* box_lo = nullptr;              // [dim][n_nodes]
    double* box_hi = nullptr;              // [dim][n_nodes]
    double* values = nullptr;              // [n] optional regression targets (original order)
    int64_t n_pad = 0;
    double bb_lo[3], bb_hi[3];
};

struct KnnView {
    int dim;
    int64_t n, n_pad, n_nodes;
    int n_levels;
    int64_t level_count[kKnnMaxLevels];
    int64_t level_offset[kKnnMaxLevels];
    const double* pts;
    const int32_t* orig;
    const double* box_lo;
    const double* box_hi;
    const double* values;
};

static inline KnnView make_view(const KnnIndex& ix) {
    KnnView v;
    v.dim = ix.dim; v.n = ix.n; v.n_pad = ix.n_pad; v.n_nodes = ix.n_nodes; v.n_levels = ix.n_levels;
    for (int i = 0; i < kKnnMaxLevels; ++i) { v.level_count[i] = ix.level_count[i]; v.level_offset[i] = ix.level_offset[i]; }
    v.pts = ix.pts; v.orig = ix.orig; v.box_lo = ix.box_lo; v.box_hi = ix.box_hi; v.values = ix.values;
    return v;
}

// Per-warp search state: lane i holds the i-th best (reduced distance, original index).
struct WarpTopK {
    double d;
    int32_t i;
};

// Exact reduced distance, operation order of sklearn's euclidean_rdist (sequential over dims, no FMA).
template <int DIM>
__device__ __forceinline__ double rdist_point(const double* q, double x, double y, double z) {
    double t = __dsub_rn(q[0], x);
    double r = __dmul_rn(t, t);
    t = __dsub_rn(q[1], y);
    r = __dadd_rn(r, __dmul_rn(t, t));
    if (DIM == 3) {
        t = __dsub_rn(q[2], z);
        r = __dadd_rn(r, __dmul_rn(t, t));
    }
    return r;
}

// Lower bound of rdist over a box; same monotone op chain as rdist_point so that
// bound(box) <= rdist(p) holds for the *computed* values of every p inside the box.
template <int DIM>
__device__ __forceinline__ double rdist_box(const double* q, const double* lo, const double* hi) {
    double r = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        double t = 0.0;
        if (q[a] < lo[a]) t = __dsub_rn(lo[a], q[a]);
        else if (q[a] > hi[a]) t = __dsub_rn(q[a], hi[a]);
        double s = __dmul_rn(t, t);
        r = (a == 0) ? s : __dadd_rn(r, s);
    }
    return r;
}

// (d, i) lexicographic "a before b"
__device__ __forceinline__ bool topk_less(double ad, int32_t ai, double bd, int32_t bi) {
    return (ad < bd) || (ad == bd && ai < bi);
}

// compare-exchange step of a warp-wide bitonic network on (d, i) pairs: partner lane ^ j, the lower lane of the pair
// keeps the smaller element iff `ascending`
__device__ __forceinline__ void topk_cmpx(double& d, int32_t& i, int j, bool ascending) {
    const double od = shfl_xor_d(d, j);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, i, j);
    const bool lower = ((threadIdx.x & 31) & j) == 0;
    const bool other_first = topk_less(od, oi, d, i);
    // keep the minimum when (lower lane, ascending) or (upper lane, descending)
    const bool keep_min = (lower == ascending);
    if (keep_min ? other_first : topk_less(d, i, od, oi)) { d = od; i = oi; }
}

// Merge the 32 candidates held one per lane (non-candidates = (+inf, INT_MAX)) into the sorted best list: bitonic sort
// of the candidates, then min(best[l], cand[31-l]) is a bitonic sequence holding the 32 smallest of the union, which
// one bitonic merge sorts again. ~210 instructions whatever the number of candidates; the one-at-a-time insertion
// costs ~27 per candidate, so this path is taken when 8 or more points of a leaf beat the current k-th distance.
__device__ __forceinline__ void topk_merge_sorted(WarpTopK& best, double cd, int32_t ci) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) topk_cmpx(cd, ci, j, (lane & k2) == 0);
    const double rd = shfl_d(cd, 31 - lane);
    const int32_t ri = __shfl_sync(0xffffffffu, ci, 31 - lane);
    if (topk_less(rd, ri, best.d, best.i)) { best.d = rd; best.i = ri; }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) topk_cmpx(best.d, best.i, j, true);
}

// One warp searches the k nearest neighbours of q. On return lane j (< k) holds the j-th neighbour
// (ascending reduced distance, ties by original index). `stack_*` is per-warp shared memory.
template <int DIM>
__device__ __forceinline__ WarpTopK warp_knn_search(const KnnView& ix, const double* q, int k,
                                                    double* stack_lb, int32_t* stack_node) {
    const int lane = threadIdx.x & 31;
    WarpTopK best;
    best.d = __longlong_as_double(0x7ff0000000000000ll);  // +inf
    best.i = 0x7fffffff;
    double kth_d = best.d;
    int32_t kth_i = best.i;

    int sp = 0;
    // stack entry: level (3 bits) | number of entries of the same push group below this one (5 bits) | node (24 bits)
    if (lane == 0) {
        stack_lb[0] = 0.0;
        stack_node[0] = (int32_t)(((uint32_t)(ix.n_levels - 1) << 29) | 0u);
    }
    sp = 1;
    __syncwarp();

    while (sp > 0) {
        --sp;
        const double lb = stack_lb[sp];
        const uint32_t enc = (uint32_t)stack_node[sp];
        __syncwarp();
        if (lb > kth_d) {
            // the children of a node were pushed nearest-on-top: everything of this group still below has a larger
            // bound and is pruned as well
            sp -= (int)((enc >> 24) & 31u);
            continue;
        }
        const int level = (int)(enc >> 29);
        const int64_t node = (int64_t)(enc & 0x00ffffffu);
        if (level == 0) {
            // leaf: 32 consecutive sorted points, one per lane
            const int64_t p = node * kKnnFan + lane;
            const bool valid = p < ix.n;
            double x = 0, y = 0, z = 0;
            int32_t oi = 0x7fffffff;
            if (valid) {
                x = ix.pts[p];
                y = ix.pts[ix.n_pad + p];
                if (DIM == 3) z = ix.pts[2 * ix.n_pad + p];
                oi = ix.orig[p];
            }
            const double rd = rdist_point<DIM>(q, x, y, z);
            bool pass = valid && (rd < kth_d || (rd == kth_d && oi < kth_i));
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (__popc(m) >= 8) {
                // many points of this leaf beat the current k-th distance (always true for the first leaves, while the
                // list is still filling): merge them in one sorting-network pass
                topk_merge_sorted(best, pass ? rd : __longlong_as_double(0x7ff0000000000000ll), pass ? oi : 0x7fffffff);
                kth_d = shfl_d(best.d, k - 1);
                kth_i = __shfl_sync(0xffffffffu, best.i, k - 1);
                m = 0;
            }
            while (m) {
                const int src = __ffs(m) - 1;
                const double cd = shfl_d(rd, src);
                const int32_t ci = __shfl_sync(0xffffffffu, oi, src);
                const bool less = (best.d < cd) || (best.d == cd && best.i < ci);
                const int pos = __popc(__ballot_sync(0xffffffffu, less));
                const double nd = shfl_up_d(best.d, 1);
                const int32_t ni = __shfl_up_sync(0xffffffffu, best.i, 1);
                if (lane > pos) { best.d = nd; best.i = ni; }
                else if (lane == pos) { best.d = cd; best.i = ci; }
                kth_d = shfl_d(best.d, k - 1);
                kth_i = __shfl_sync(0xffffffffu, best.i, k - 1);
                if (lane == src) pass = false;
                pass = pass && (rd < kth_d || (rd == kth_d && oi < kth_i));
                m = __ballot_sync(0xffffffffu, pass);
            }
        } else {
            // internal node: lane c looks at child c (a node of level-1)
            const int clevel = level - 1;
            const int64_t child = node * kKnnFan + lane;
            const bool exists = child < ix.level_count[clevel];
            double clb = __longlong_as_double(0x7ff0000000000000ll);
            if (exists) {
                const int64_t o = ix.level_offset[clevel] + child;
                double lo[DIM], hi[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    lo[a] = ix.box_lo[(int64_t)a * ix.n_nodes + o];
                    hi[a] = ix.box_hi[(int64_t)a * ix.n_nodes + o];
                }
                clb = rdist_box<DIM>(q, lo, hi);
            }
            const bool push = exists && !(clb > kth_d);
            const uint32_t pm = __ballot_sync(0xffffffffu, push);
            const int npush = __popc(pm);
            // rank among pushed children by (lb, lane) ascending
            int rank = 0;
            uint32_t mm = pm;
            while (mm) {
                const int j = __ffs(mm) - 1;
                mm &= mm - 1;
                const double v = shfl_d(clb, j);
                rank += (v < clb || (v == clb && j < lane)) ? 1 : 0;
            }
            if (push) {
                const int below = npush - 1 - rank;        // nearest child ends on top
                stack_lb[sp + below] = clb;
                stack_node[sp + below] = (int32_t)(((uint32_t)clevel << 29) | ((uint32_t)below << 24) | (uint32_t)child);
            }
            sp += npush;
            __syncwarp();
        }
    }
    return best;
}

// numpy's pairwise summation for n <= 128 contiguous doubles (the order np.sum(axis=1) uses for the
// k weights of one query row; sklearn/neighbors/_regression.py: num = np.sum(y * w, axis=1)).
// Lane j holds a[j] for j < n. Result valid in lane 0.
__device__ __forceinline__ double warp_numpy_pairwise_sum(double a, int n) {
    const int lane = threadIdx.x & 31;
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, shfl_d(a, i));
        return res;
    }
    const int n8 = n - (n & 7);
    double r = a;  // lanes 0..7: r[j] = a[j]
    for (int i = 8; i < n8; i += 8) {
        double v = shfl_d(a, (lane & 7) + i);
        r = __dadd_rn(r, v);
    }
    // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))
    r = __dadd_rn(r, shfl_xor_d(r, 1));
    r = __dadd_rn(r, shfl_xor_d(r, 2));
    r = __dadd_rn(r, shfl_xor_d(r, 4));
    double res = shfl_d(r, 0);
    for (int i = n8; i < n; ++i) res = __dadd_rn(res, shfl_d(a, i));
    return res;
}

// Inverse-distance prediction of sklearn's KNeighborsRegressor(weights="distance")
// (sklearn/neighbors/_base.py _get_weights + _regression.py predict), bit-for-bit:
// w = 1/dist, rows containing a zero distance use the 0/1 indicator of the zeros instead.
__device__ __forceinline__ double warp_idw_predict(const WarpTopK& best, int k, const double* values) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < k;
    double dist = act ? __dsqrt_rn(best.d) : 1.0;
    const uint32_t zero_mask = __ballot_sync(0xffffffffu, act && dist == 0.0);
    double w;
    if (zero_mask) w = (act && dist == 0.0) ? 1.0 : 0.0;
    else w = __ddiv_rn(1.0, dist);
    double y = act ? values[best.i] : 0.0;
    double p = __dmul_rn(y, w);
    if (!act) { p = 0.0; w = 0.0; }
    const double num = warp_numpy_pairwise_sum(p, k);
    const double den = warp_numpy_pairwise_sum(w, k);
    return __ddiv_rn(num, den);
}

}  // namespace s3
