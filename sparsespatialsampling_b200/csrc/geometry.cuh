// Point-in-shape tests of the geometry plugins (device side of GeometryObject.check_cell).
//
// Each test restates the torch expression of the reference class in fp64 with the SAME rounding
// sequence torch's CPU kernels use (probed on the build host, see DESIGN.md "floating-point order"):
//   * tensor.norm(dim=1) / tensor.norm()      -> sqrt of a fused-multiply-add chain  acc = fma(x, x, acc)
//   * torch.cross                             -> c0 = fma(a1, b2, -(a2*b1)), ...
//   * (a * b).sum(-1) over 3 components       -> separate products, sequential adds
//   * torch.dot of two 3-vectors              -> (x0*y0 + x2*y2) + x1*y1, no fma
//   * element-wise expressions                -> one rounding per operator, no contraction
// Reference files: sparseSpatialSampling/geometry/{cube,sphere,cylinder,triangle,prism,tetrahedron,
// pyramid}_geometry.py, geometry_STL_3d.py, coordinates_2d.py.
#pragma once
#include "common.cuh"

namespace s3 {

enum GeomType : int {
    GEOM_CUBE = 0,
    GEOM_SPHERE = 1,
    GEOM_CYLINDER = 2,
    GEOM_TRIANGLE = 3,
    GEOM_PRISM = 4,
    GEOM_TETRA = 5,
    GEOM_PYRAMID = 6,
    GEOM_STL = 7,
    GEOM_POLY2D = 8,
    GEOM_CUSTOM = 9,
};

// header: int32 [G][4] = {type, keep_inside, param offset (doubles), n_extra}
struct GeomHdr {
    int type, keep_inside, offset, n_extra;
};

__device__ __forceinline__ double norm_fma(const double* v, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = __fma_rn(v[i], v[i], acc);
    return __dsqrt_rn(acc);
}

__device__ __forceinline__ void cross_torch(const double* a, const double* b, double* c) {
    c[0] = __fma_rn(a[1], b[2], -__dmul_rn(a[2], b[1]));
    c[1] = __fma_rn(a[2], b[0], -__dmul_rn(a[0], b[2]));
    c[2] = __fma_rn(a[0], b[1], -__dmul_rn(a[1], b[0]));
}

__device__ __forceinline__ double dot3_blas(const double* x, const double* y) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x[0], y[0]), __dmul_rn(x[2], y[2])), __dmul_rn(x[1], y[1]));
}

// cube_geometry.py:71 (flowtorch mask_box, inclusive bounds)
__device__ __forceinline__ bool in_cube(const double* p, const double* par, int dim) {
    bool in = true;
    for (int a = 0; a < dim; ++a) in = in && (p[a] >= par[a]) && (p[a] <= par[dim + a]);
    return in;
}

// sphere_geometry.py:69 (flowtorch mask_sphere: ||v - c|| <= r)
__device__ __forceinline__ bool in_sphere(const double* p, const double* par, int dim) {
    double v[3];
    for (int a = 0; a < dim; ++a) v[a] = __dsub_rn(p[a], par[a]);
    return norm_fma(v, dim) <= par[dim];
}

// cylinder_geometry.py:126-157; par = p0[3], axis[3], norm, r0, r1, cone
__device__ __forceinline__ bool in_cylinder(const double* p, const double* par) {
    double dv[3], cr[3];
    for (int a = 0; a < 3; ++a) dv[a] = __dsub_rn(p[a], par[a]);
    const double* axis = par + 3;
    const double nrm = par[6];
    cross_torch(axis, dv, cr);
    const double nd = __ddiv_rn(norm_fma(cr, 3), nrm);
    double s = __dmul_rn(dv[0], axis[0]);
    s = __dadd_rn(s, __dmul_rn(dv[1], axis[1]));
    s = __dadd_rn(s, __dmul_rn(dv[2], axis[2]));
    const double proj = __ddiv_rn(s, nrm);
    const bool within_h = (0.0 <= proj) && (proj <= nrm);
    double rad = par[7];
    if (par[9] != 0.0) rad = __dadd_rn(par[7], __dmul_rn(__ddiv_rn(proj, nrm), __dsub_rn(par[8], par[7])));
    return within_h && (nd <= rad);
}

// triangle_geometry.py:80-104; tri = P0(2), P1(2), P2(2); point (x, y)
__device__ __forceinline__ bool in_triangle(double x, double y, const double* t) {
    const double *P0 = t, *P1 = t + 2, *P2 = t + 4;
    // d1 = (P1-P0) x (v-P0)
    double ax = __dsub_rn(P1[0], P0[0]), ay = __dsub_rn(P1[1], P0[1]);
    double bx = __dsub_rn(x, P0[0]), by = __dsub_rn(y, P0[1]);
    const double d1 = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    // d2 = (P2-P1) x (v-P1)
    ax = __dsub_rn(P2[0], P1[0]); ay = __dsub_rn(P2[1], P1[1]);
    double cx = __dsub_rn(x, P1[0]), cy = __dsub_rn(y, P1[1]);
    const double d2 = __dsub_rn(__dmul_rn(ax, cy), __dmul_rn(ay, cx));
    // d3 = (P0-P2) x (v-P0)   (sic: measured from P0, triangle_geometry.py:98)
    ax = __dsub_rn(P0[0], P2[0]); ay = __dsub_rn(P0[1], P2[1]);
    const double d3 = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    const bool neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
    const bool pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
    return !(neg && pos);
}

// prism_geometry.py:90-118; par = p0[3], axis[3], norm, dimA, dimB, tri[6]
__device__ __forceinline__ bool in_prism(const double* p, const double* par) {
    double dv[3];
    for (int a = 0; a < 3; ++a) dv[a] = __dsub_rn(p[a], par[a]);
    const double* axis = par + 3;
    const double nrm = par[6];
    double s = __dmul_rn(dv[0], axis[0]);
    s = __dadd_rn(s, __dmul_rn(dv[1], axis[1]));
    s = __dadd_rn(s, __dmul_rn(dv[2], axis[2]));
    const double proj = __ddiv_rn(s, nrm);
    const bool within_h = (0.0 <= proj) && (proj <= nrm);
    const int da = (int)par[7], db = (int)par[8];
    return within_h && in_triangle(p[da], p[db], par + 9);
}

// tetrahedron_geometry.py:121-140; par = positions[4][3], normals[4][3] (normal of point p at par+12+3p)
__device__ __forceinline__ bool in_tetra(const double* p, const double* par) {
    bool outside = false;
    for (int q = 0; q < 4; ++q) {
        double v[3];
        for (int a = 0; a < 3; ++a) v[a] = __dsub_rn(p[a], par[3 * q + a]);
        outside = outside || (dot3_blas(v, par + 12 + 3 * q) < 0.0);
    }
    return !outside;
}

// pyramid_geometry.py:156-170: union of two tetrahedra
__device__ __forceinline__ bool in_pyramid(const double* p, const double* par) {
    return in_tetra(p, par) || in_tetra(p, par + 24);
}

// Closed triangulated surface (geometry_STL_3d.py:81-103, pyvista select_enclosed_points with
// check_surface=False).  VTK is not available offline, so this is the documented restatement:
// a point is inside if it lies within `tol` (absolute, = tolerance * bbox diagonal as in
// vtkSelectEnclosedPoints) of the surface, else by the parity of +x ray crossings; the ray is cast
// through y/z perturbed by irrational offsets of relative size 1e-9 so edges/vertices are rarely hit, and
// a ray that still passes exactly through an edge or a vertex of the projected mesh is counted once per
// surface crossing by the half-open rule of stl_edge_side (below).
// par = lo[3], hi[3], tol, then n_tri * 9 doubles.
__device__ __forceinline__ double pt_tri_dist2(const double* p, const double* a, const double* b, const double* c) {
    // closest point on triangle (Ericson, Real-Time Collision Detection 5.1.5)
    double ab[3], ac[3], ap[3];
    for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; }
    const double d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2];
    const double d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
    double q[3];
    if (d1 <= 0 && d2 <= 0) { for (int i = 0; i < 3; ++i) q[i] = a[i]; }
    else {
        double bp[3];
        for (int i = 0; i < 3; ++i) bp[i] = p[i] - b[i];
        const double d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2];
        const double d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
        if (d3 >= 0 && d4 <= d3) { for (int i = 0; i < 3; ++i) q[i] = b[i]; }
        else {
            const double vc = d1 * d4 - d3 * d2;
            if (vc <= 0 && d1 >= 0 && d3 <= 0) {
                const double v = d1 / (d1 - d3);
                for (int i = 0; i < 3; ++i) q[i] = a[i] + v * ab[i];
            } else {
                double cp[3];
                for (int i = 0; i < 3; ++i) cp[i] = p[i] - c[i];
                const double d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2];
                const double d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
                if (d6 >= 0 && d5 <= d6) { for (int i = 0; i < 3; ++i) q[i] = c[i]; }
                else {
                    const double vb = d5 * d2 - d1 * d6;
                    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
                        const double w = d2 / (d2 - d6);
                        for (int i = 0; i < 3; ++i) q[i] = a[i] + w * ac[i];
                    } else {
                        const double va = d3 * d6 - d5 * d4;
                        if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
                            const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            for (int i = 0; i < 3; ++i) q[i] = b[i] + w * (c[i] - b[i]);
                        } else {
                            const double den = 1.0 / (va + vb + vc);
                            const double v = vb * den, w = vc * den;
                            for (int i = 0; i < 3; ++i) q[i] = a[i] + ab[i] * v + ac[i] * w;
                        }
                    }
                }
            }
        }
    }
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return dx * dx + dy * dy + dz * dz;
}

// Side of the projected ray point (the origin; u, v are edge end points relative to it, in the y-z plane) with
// respect to the directed edge u -> v: the sign of cross(u, v), evaluated WITHOUT fused multiply-add so that the two
// triangles sharing an edge (which see it as u -> v and v -> u) get exactly opposite values. A zero is resolved as if
// the point were shifted by (eps, eps^2) in (y, z): d cross / dy = u.z - v.z, d cross / dz = v.y - u.y. This is the
// usual top-left fill rule; it is antisymmetric in (u, v) like the value itself, hence independent of the winding of
// the triangles, and a ray through a shared edge or vertex of a closed surface is claimed by exactly one triangle
// per crossing (edge-on triangles, whose three sides cannot agree, by none).
__device__ __forceinline__ int stl_edge_side(double uy, double uz, double vy, double vz) {
    const double f = __dsub_rn(__dmul_rn(uy, vz), __dmul_rn(uz, vy));
    if (f > 0) return 1;
    if (f < 0) return -1;
    if (uz != vz) return uz > vz ? 1 : -1;
    if (uy != vy) return vy > uy ? 1 : -1;
    return 0;
}

// +x ray from (px, py, pz) against triangle (a, b, c): point-in-triangle in the y-z plane, then the x of the hit
__device__ __forceinline__ bool stl_ray_crosses(double px, double py, double pz, const double* a, const double* b,
                                                const double* c) {
    const double ay = a[1] - py, az = a[2] - pz, by = b[1] - py, bz = b[2] - pz, cy = c[1] - py, cz = c[2] - pz;
    const int e0 = stl_edge_side(ay, az, by, bz), e1 = stl_edge_side(by, bz, cy, cz), e2 = stl_edge_side(cy, cz, ay, az);
    if (!((e0 > 0 && e1 > 0 && e2 > 0) || (e0 < 0 && e1 < 0 && e2 < 0))) return false;
    const double s0 = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by));
    const double s1 = __dsub_rn(__dmul_rn(by, cz), __dmul_rn(bz, cy));
    const double s2 = __dsub_rn(__dmul_rn(cy, az), __dmul_rn(cz, ay));
    const double sum = __dadd_rn(__dadd_rn(s0, s1), s2);
    if (sum == 0.0) return false;
    const double num = __dadd_rn(__dadd_rn(__dmul_rn(s1, a[0]), __dmul_rn(s2, b[0])), __dmul_rn(s0, c[0]));
    return num / sum > px;
}

__device__ __forceinline__ bool in_stl(const double* p, const double* par, int n_tri) {
    const double* lo = par;
    const double* hi = par + 3;
    const double tol = par[6];
    for (int a = 0; a < 3; ++a)
        if (p[a] < lo[a] - tol || p[a] > hi[a] + tol) return false;
    const double* tri = par + 7;
    const double ext = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const double py = p[1] + ext * 1.4142135623730951e-9;
    const double pz = p[2] + ext * 1.7320508075688772e-9;
    int crossings = 0;
    bool near = false;
    for (int t = 0; t < n_tri; ++t) {
        const double* a = tri + 9 * t;
        const double* b = a + 3;
        const double* c = a + 6;
        if (pt_tri_dist2(p, a, b, c) <= tol * tol) near = true;
        if (stl_ray_crosses(p[0], py, pz, a, b, c)) ++crossings;
    }
    return near || (crossings & 1);
}

// Closed polygon (coordinates_2d.py:54-73, shapely Point.within == strict interior): points on the
// boundary are outside. Even-odd crossing rule; par = lo[2], hi[2], then n_v * 2 doubles.
__device__ __forceinline__ bool in_poly2d(const double* p, const double* par, int n_v) {
    const double* v = par + 4;
    const double x = p[0], y = p[1];
    if (x < par[0] || x > par[2] || y < par[1] || y > par[3]) return false;
    bool inside = false;
    for (int i = 0, j = n_v - 1; i < n_v; j = i++) {
        const double xi = v[2 * i], yi = v[2 * i + 1], xj = v[2 * j], yj = v[2 * j + 1];
        // on-segment test (exact orientation sign + bounding interval) -> boundary -> outside
        const double cr = (xj - xi) * (y - yi) - (yj - yi) * (x - xi);
        if (cr == 0.0 && x >= fmin(xi, xj) && x <= fmax(xi, xj) && y >= fmin(yi, yj) && y <= fmax(yi, yj)) return false;
        if ((yi > y) != (yj > y)) {
            const double xint = (xj - xi) * (y - yi) / (yj - yi) + xi;
            if (x < xint) inside = !inside;
        }
    }
    return inside;
}

__device__ __forceinline__ bool point_in_geometry(const GeomHdr& h, const double* par, const double* p, int dim) {
    switch (h.type) {
        case GEOM_CUBE: return in_cube(p, par, dim);
        case GEOM_SPHERE: return in_sphere(p, par, dim);
        case GEOM_CYLINDER: return in_cylinder(p, par);
        case GEOM_TRIANGLE: return in_triangle(p[0], p[1], par);
        case GEOM_PRISM: return in_prism(p, par);
        case GEOM_TETRA: return in_tetra(p, par);
        case GEOM_PYRAMID: return in_pyramid(p, par);
        case GEOM_STL: return in_stl(p, par, h.n_extra);
        case GEOM_POLY2D: return in_poly2d(p, par, h.n_extra);
        default: return false;
    }
}

// GeometryObject._apply_mask (geometry_base.py:40-76): n_in of n_nodes nodes are inside.
__device__ __forceinline__ bool apply_mask(int n_in, int n_nodes, bool keep_inside, bool refine_mode) {
    if (!refine_mode) return keep_inside ? (n_in == 0) : (n_in == n_nodes);
    return keep_inside ? (n_in != n_nodes) : (n_in > 0);
}

}  // namespace s3
