/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never on the product path.
 *
 * Exact brute-force restatement of the neighbour search the reference delegates to scikit-learn 1.9.0
 * (third-party, not vendored in the reference; call sites sparseSpatialSampling/s_cube.py:161-163,224,328,372
 * and sparseSpatialSampling/export.py:120,423-441).  sklearn's KD-tree ranks candidates by the reduced
 * distance  sum_j (q_j - x_j)^2  accumulated dimension by dimension in double precision without fused
 * multiply-add (sklearn/metrics/_dist_metrics.pyx euclidean_rdist) and reports sqrt() of it; a brute-force
 * scan over the same values therefore returns the same neighbours whenever there is no exact tie
 * (ties: smaller index first here; unspecified in sklearn).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void s3o_knn(const double* X, int64_t n, int d, const double* Q, int64_t nq, int k, int64_t* idx, double* dist) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        double* bd = dist + qi * k;
        int64_t* bi = idx + qi * k;
        int filled = 0;
        const double* q = Q + qi * d;
        for (int64_t i = 0; i < n; ++i) {
            const double* x = X + i * d;
            double r = 0.0;
            for (int j = 0; j < d; ++j) {
                double t = q[j] - x[j];
                r += t * t;
            }
            if (filled == k && !(r < bd[k - 1])) continue; /* equal to the k-th: later index loses */
            int pos = filled < k ? filled : k - 1;
            while (pos > 0 && bd[pos - 1] > r) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = r;
            bi[pos] = i;
            if (filled < k) ++filled;
        }
        for (int j = 0; j < k; ++j) bd[j] = sqrt(bd[j]);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Point-in-shape tests of the reference's geometry classes (sparseSpatialSampling/geometry/*.py),
 * restated with the rounding sequence of the torch CPU operators they call (probed against torch
 * 2.11 on the build host): norm() = sqrt of an fma chain, cross() = fma(a1,b2,-(a2*b1)),
 * 3-term (a*b).sum(-1) = separate products + sequential adds, dot() of 3-vectors = (x0y0 + x2y2) + x1y1.
 * Parameter layout = the blocks built by sparsespatialsampling_b200/geometry (type ids below).
 * ------------------------------------------------------------------------------------------------ */
enum { G_CUBE = 0, G_SPHERE, G_CYLINDER, G_TRIANGLE, G_PRISM, G_TETRA, G_PYRAMID, G_STL, G_POLY2D };

static double norm_fma(const double* v, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = fma(v[i], v[i], acc);
    return sqrt(acc);
}

static int o_cube(const double* p, const double* par, int dim) { /* cube_geometry.py:71, inclusive */
    for (int a = 0; a < dim; ++a)
        if (!(p[a] >= par[a] && p[a] <= par[dim + a])) return 0;
    return 1;
}

static int o_sphere(const double* p, const double* par, int dim) { /* sphere_geometry.py:69 */
    double v[3];
    for (int a = 0; a < dim; ++a) v[a] = p[a] - par[a];
    return norm_fma(v, dim) <= par[dim];
}

static int o_cylinder(const double* p, const double* par) { /* cylinder_geometry.py:126-157 */
    double dv[3], cr[3];
    const double* ax = par + 3;
    const double nrm = par[6];
    for (int a = 0; a < 3; ++a) dv[a] = p[a] - par[a];
    cr[0] = fma(ax[1], dv[2], -(ax[2] * dv[1]));
    cr[1] = fma(ax[2], dv[0], -(ax[0] * dv[2]));
    cr[2] = fma(ax[0], dv[1], -(ax[1] * dv[0]));
    const double nd = norm_fma(cr, 3) / nrm;
    double s = dv[0] * ax[0];
    s = s + dv[1] * ax[1];
    s = s + dv[2] * ax[2];
    const double proj = s / nrm;
    double rad = par[7];
    if (par[9] != 0.0) rad = par[7] + (proj / nrm) * (par[8] - par[7]);
    return (0.0 <= proj) && (proj <= nrm) && (nd <= rad);
}

static int o_triangle(double x, double y, const double* t) { /* triangle_geometry.py:80-104 */
    const double *P0 = t, *P1 = t + 2, *P2 = t + 4;
    const double bx = x - P0[0], by = y - P0[1], cx = x - P1[0], cy = y - P1[1];
    const double d1 = (P1[0] - P0[0]) * by - (P1[1] - P0[1]) * bx;
    const double d2 = (P2[0] - P1[0]) * cy - (P2[1] - P1[1]) * cx;
    const double d3 = (P0[0] - P2[0]) * by - (P0[1] - P2[1]) * bx;
    const int neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
    const int pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
    return !(neg && pos);
}

static int o_prism(const double* p, const double* par) { /* prism_geometry.py:90-118 */
    double dv[3];
    const double* ax = par + 3;
    const double nrm = par[6];
    for (int a = 0; a < 3; ++a) dv[a] = p[a] - par[a];
    double s = dv[0] * ax[0];
    s = s + dv[1] * ax[1];
    s = s + dv[2] * ax[2];
    const double proj = s / nrm;
    return (0.0 <= proj) && (proj <= nrm) && o_triangle(p[(int)par[7]], p[(int)par[8]], par + 9);
}

static int o_tetra(const double* p, const double* par) { /* tetrahedron_geometry.py:121-140 */
    for (int q = 0; q < 4; ++q) {
        double v[3];
        const double* n = par + 12 + 3 * q;
        for (int a = 0; a < 3; ++a) v[a] = p[a] - par[3 * q + a];
        const double dot = (v[0] * n[0] + v[2] * n[2]) + v[1] * n[1];
        if (dot < 0.0) return 0;
    }
    return 1;
}

static double o_pt_tri_dist2(const double* p, const double* a, const double* b, const double* c) {
    double ab[3], ac[3], ap[3], bp[3], cp[3], q[3];
    for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; bp[i] = p[i] - b[i]; cp[i] = p[i] - c[i]; }
    const double d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2];
    const double d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
    const double d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2];
    const double d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
    const double d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2];
    const double d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
    const double vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    if (d1 <= 0 && d2 <= 0) { for (int i = 0; i < 3; ++i) q[i] = a[i]; }
    else if (d3 >= 0 && d4 <= d3) { for (int i = 0; i < 3; ++i) q[i] = b[i]; }
    else if (vc <= 0 && d1 >= 0 && d3 <= 0) { const double v = d1 / (d1 - d3); for (int i = 0; i < 3; ++i) q[i] = a[i] + v * ab[i]; }
    else if (d6 >= 0 && d5 <= d6) { for (int i = 0; i < 3; ++i) q[i] = c[i]; }
    else if (vb <= 0 && d2 >= 0 && d6 <= 0) { const double w = d2 / (d2 - d6); for (int i = 0; i < 3; ++i) q[i] = a[i] + w * ac[i]; }
    else if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); for (int i = 0; i < 3; ++i) q[i] = b[i] + w * (c[i] - b[i]); }
    else { const double den = 1.0 / (va + vb + vc); const double v = vb * den, w = vc * den; for (int i = 0; i < 3; ++i) q[i] = a[i] + ab[i] * v + ac[i] * w; }
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return dx * dx + dy * dy + dz * dz;
}

/* geometry_STL_3d.py:81-103 (pyvista select_enclosed_points, check_surface=False) -- VTK is not available
 * offline; documented restatement: within tol of the surface => inside, else parity of +x ray crossings
 * (ray nudged in y/z; a ray that still runs exactly through an edge or vertex of the projected mesh is resolved
 * by the top-left rule below, i.e. as if nudged further by (eps, eps^2)). par = lo[3], hi[3], tol, triangles[n][9].
 * PARITY UNPINNED beyond the reference's own tests (tests/test_geometry_STL.py on tests/cube.stl) and the
 * closed-form solids of tests/test_geometry_surfaces_*.py. This file is compiled with -ffp-contract=off: the sign
 * of an edge function has to be the exact negative for the neighbouring triangle. */
static int o_edge_side(double uy, double uz, double vy, double vz) {
    const double m0 = uy * vz, m1 = uz * vy;
    const double f = m0 - m1;
    if (f > 0) return 1;
    if (f < 0) return -1;
    if (uz != vz) return uz > vz ? 1 : -1;
    if (uy != vy) return vy > uy ? 1 : -1;
    return 0;
}
static int o_ray_crosses(double px, double py, double pz, const double* a, const double* b, const double* c) {
    const double ay = a[1] - py, az = a[2] - pz, by = b[1] - py, bz = b[2] - pz, cy = c[1] - py, cz = c[2] - pz;
    const int e0 = o_edge_side(ay, az, by, bz), e1 = o_edge_side(by, bz, cy, cz), e2 = o_edge_side(cy, cz, ay, az);
    if (!((e0 > 0 && e1 > 0 && e2 > 0) || (e0 < 0 && e1 < 0 && e2 < 0))) return 0;
    const double s0 = ay * bz - az * by, s1 = by * cz - bz * cy, s2 = cy * az - cz * ay;
    const double sum = (s0 + s1) + s2;
    if (sum == 0.0) return 0;
    const double num = (s1 * a[0] + s2 * b[0]) + s0 * c[0];
    return num / sum > px;
}
static int o_stl(const double* p, const double* par, int n_tri) {
    const double *lo = par, *hi = par + 3, tol = par[6], *tri = par + 7;
    for (int a = 0; a < 3; ++a)
        if (p[a] < lo[a] - tol || p[a] > hi[a] + tol) return 0;
    const double ext = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const double py = p[1] + ext * 1.4142135623730951e-9, pz = p[2] + ext * 1.7320508075688772e-9;
    int crossings = 0, near = 0;
    for (int t = 0; t < n_tri; ++t) {
        const double *a = tri + 9 * t, *b = a + 3, *c = a + 6;
        if (o_pt_tri_dist2(p, a, b, c) <= tol * tol) near = 1;
        crossings += o_ray_crosses(p[0], py, pz, a, b, c);
    }
    return near || (crossings & 1);
}

/* coordinates_2d.py:54-73 (shapely Point.within: strict interior). par = lo[2], hi[2], vertices[n][2].
 * PARITY UNPINNED beyond tests/test_coordinates_2d_geometry.py. */
static int o_poly2d(const double* p, const double* par, int nv) {
    const double* v = par + 4;
    const double x = p[0], y = p[1];
    if (x < par[0] || x > par[2] || y < par[1] || y > par[3]) return 0;
    int inside = 0;
    for (int i = 0, j = nv - 1; i < nv; j = i++) {
        const double xi = v[2 * i], yi = v[2 * i + 1], xj = v[2 * j], yj = v[2 * j + 1];
        const double cr = (xj - xi) * (y - yi) - (yj - yi) * (x - xi);
        if (cr == 0.0 && x >= fmin(xi, xj) && x <= fmax(xi, xj) && y >= fmin(yi, yj) && y <= fmax(yi, yj)) return 0;
        if ((yi > y) != (yj > y)) {
            const double xint = (xj - xi) * (y - yi) / (yj - yi) + xi;
            if (x < xint) inside = !inside;
        }
    }
    return inside;
}

void s3o_points_inside(int type, const double* par, int n_extra, const double* pts, int64_t n, int dim, uint8_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        const double* p = pts + i * dim;
        int r = 0;
        switch (type) {
            case G_CUBE: r = o_cube(p, par, dim); break;
            case G_SPHERE: r = o_sphere(p, par, dim); break;
            case G_CYLINDER: r = o_cylinder(p, par); break;
            case G_TRIANGLE: r = o_triangle(p[0], p[1], par); break;
            case G_PRISM: r = o_prism(p, par); break;
            case G_TETRA: r = o_tetra(p, par); break;
            case G_PYRAMID: r = o_tetra(p, par) || o_tetra(p, par + 24); break;
            case G_STL: r = o_stl(p, par, n_extra); break;
            case G_POLY2D: r = o_poly2d(p, par, n_extra); break;
            default: r = 0;
        }
        out[i] = (uint8_t)r;
    }
}
