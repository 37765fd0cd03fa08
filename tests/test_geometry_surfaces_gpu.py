"""
GPU tests of the surface-defined geometry plugins (GeometrySTL3D, GeometryCoordinates2D): the reference's own unit
tests (sparseSpatialSampling/tests/test_geometry_STL.py:31-63, test_coordinates_2d_geometry.py:37-71, restated with a
generated unit-cube STL), device-vs-oracle parity on dense point sets, and analytic sanity checks.
Beyond the reference's unit tests the parity of these two masks with VTK / shapely is unpinned.
"""
import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc
from tests.stl_util import write_binary_stl, cube_triangles, icosphere_triangles
from tests.test_geometry_gpu import DummyCells

pytestmark = pytest.mark.gpu


@pytest.fixture
def cube_stl(tmp_path):
    p = tmp_path / "cube.stl"
    write_binary_stl(p, cube_triangles())
    return str(p)


@pytest.mark.parametrize("keep_inside,expected", [(False, (False, True, False)), (True, (True, False, False))])
def test_stl_reference_unit_triples(cuda, cube_stl, keep_inside, expected):
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    g = GeometrySTL3D("cube", keep_inside=keep_inside, path_stl_file=cube_stl)
    c = DummyCells()
    # inside cell = the cube's own corners: on-surface points count as inside
    got = (g.check_cell(c.cell_outside_3D), g.check_cell(c.cell_inside_3D), g.check_cell(c.cell_partially_3D))
    assert got == expected
    assert g.type == "STL" and g.main_width == 1.0 and pt.allclose(g.center, pt.full((3,), 0.5, dtype=pt.float64))


def test_stl_pre_check_cell(cuda, cube_stl):
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    g = GeometrySTL3D("cube", keep_inside=False, path_stl_file=cube_stl)
    c = DummyCells()
    assert g.pre_check_cell(c.cell_inside_3D) is True
    assert g.pre_check_cell(c.cell_outside_3D) is False


def test_stl_rejects_open_surfaces_and_decimation(cuda, tmp_path, cube_stl):
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    p = tmp_path / "open.stl"
    write_binary_stl(p, cube_triangles()[:-2])            # one face missing
    with pytest.raises(RuntimeError):
        GeometrySTL3D("open", False, str(p))
    with pytest.raises(NotImplementedError):
        GeometrySTL3D("cube", False, cube_stl, reduce_by=0.5)


def test_stl_sphere_matches_oracle_and_analytic(cuda, tmp_path):
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    from sparsespatialsampling_b200.geometry.device import nodes_inside
    p = tmp_path / "sphere.stl"
    write_binary_stl(p, icosphere_triangles(3, radius=0.7, center=(0.1, -0.2, 0.3)))
    g = GeometrySTL3D("sphere", False, str(p))
    assert g._triangles.shape[0] == 1280
    rng = np.random.default_rng(0)
    pts = rng.random((20000, 3)) * 2.0 - 1.0 + np.array([0.1, -0.2, 0.3])
    got = nodes_inside(g, pt.from_numpy(pts)).numpy()
    assert np.array_equal(got, orc.points_inside(g, pts))
    r = np.linalg.norm(pts - np.array([0.1, -0.2, 0.3]), axis=1)
    clear = np.abs(r - 0.7) > 0.03                       # away from the faceted surface / tolerance band
    assert np.array_equal(got[clear], (r < 0.7)[clear])


@pytest.mark.parametrize("keep_inside,expected", [(False, (False, True, False)), (True, (True, False, False))])
def test_polygon_reference_unit_triples(cuda, keep_inside, expected):
    from sparsespatialsampling_b200.geometry import GeometryCoordinates2D
    g = GeometryCoordinates2D("square", keep_inside=keep_inside,
                              coordinates=[(-1, -1), (-1, 1.25), (1.25, 1.25), (1.25, -1)])
    c = DummyCells()
    got = (g.check_cell(c.cell_outside_2D), g.check_cell(c.cell_inside_2D), g.check_cell(c.cell_partially_2D))
    assert got == expected
    assert g.type == "coord_2D" and g.main_width == 2.25


def test_polygon_pre_check_and_boundary(cuda):
    from sparsespatialsampling_b200.geometry import GeometryCoordinates2D
    from sparsespatialsampling_b200.geometry.device import nodes_inside
    g = GeometryCoordinates2D("square", False, [(-1, -1), (-1, 1.25), (1.25, 1.25), (1.25, -1)])
    assert g.pre_check_cell(pt.tensor([[0.0, 0.0]])) is True
    assert g.pre_check_cell(pt.tensor([[5.0, 5.0]])) is False
    # shapely's within() is the strict interior: boundary points are outside
    edge = pt.tensor([[-1.0, 0.0], [1.25, 1.25], [0.0, 1.25], [0.0, 0.0]], dtype=pt.float64)
    assert nodes_inside(g, edge).tolist() == [False, False, False, True]


def test_polygon_star_matches_oracle(cuda):
    from sparsespatialsampling_b200.geometry import GeometryCoordinates2D
    from sparsespatialsampling_b200.geometry.device import nodes_inside
    ang = np.linspace(0, 2 * np.pi, 21)[:-1]
    rad = np.where(np.arange(20) % 2 == 0, 1.0, 0.45)
    star = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    g = GeometryCoordinates2D("star", False, star.tolist() + [star[0].tolist()])      # closed ring input
    assert g._vertices.shape[0] == 20
    pts = np.random.default_rng(1).random((30000, 2)) * 2.4 - 1.2
    got = nodes_inside(g, pt.from_numpy(pts)).numpy()
    assert np.array_equal(got, orc.points_inside(g, pts))
    assert 0.2 < got.mean() < 0.4


def test_grid_generation_with_stl_and_polygon_matches_oracle(cuda, tmp_path):
    import synth
    import sparsespatialsampling_b200.geometry as geo
    from sparsespatialsampling_b200.s_cube import SamplingTree
    # 3-D: STL sphere body inside a box domain
    p = tmp_path / "ball.stl"
    write_binary_stl(p, icosphere_triangles(2, radius=0.3, center=(0.8, 1.0, 0.16)))
    x = synth.cylinder3d_cloud(5000, seed=4)
    m = synth.wake_metric(x, xc=0.8, yc=1.0)
    mk = lambda: [geo.CubeGeometry("domain", True, synth.CYL3D["lower"], synth.CYL3D["upper"]),
                  geo.GeometrySTL3D("ball", False, str(p), refine=True)]
    kw = dict(uniform_level=3, min_metric=0.4, n_cells_iter_start=10)
    tree = SamplingTree(x, m, mk(), **kw, sdm_order=1)
    tree.refine()
    o = orc.OracleTree(x.numpy(), m.numpy(), mk(), **kw, sdm_order=1).refine()
    assert list(tree._leaf_cells) == o.leaf_order
    assert np.array_equal(tree.all_centers.numpy(), o.all_centers)
    # 2-D: polygon body (the OAT15-style set-up of the reference's example uses GeometryCoordinates2D)
    x2 = synth.airfoil2d_cloud(5000, seed=6)
    m2 = synth.wake_metric(x2, xc=1.0, yc=0.0)
    wedge = [(0.0, 0.0), (1.0, 0.06), (1.0, -0.06)]
    mk2 = lambda: [geo.CubeGeometry("domain", True, synth.AIRFOIL2D["lower"], synth.AIRFOIL2D["upper"]),
                   geo.GeometryCoordinates2D("airfoil", False, wedge, refine=True, min_refinement_level=7)]
    kw2 = dict(uniform_level=4, n_cells=1200, n_cells_iter_start=40)
    t2 = SamplingTree(x2, m2, mk2(), **kw2, sdm_order=1)
    t2.refine()
    o2 = orc.OracleTree(x2.numpy(), m2.numpy(), mk2(), **kw2, sdm_order=1).refine()
    assert list(t2._leaf_cells) == o2.leaf_order
    assert np.array_equal(t2.all_centers.numpy(), o2.all_centers)
    assert np.array_equal(t2.all_levels.numpy(), o2.all_levels)


# ---------------------------------------------------------------------------------------------------------------
# Closed-form solids. The STL rule restated from geometry_STL_3d.py:81-103 / vtkSelectEnclosedPoints:
# a point within tol = 0.001 * bounding-box diagonal of the surface counts as inside, everything else by ray parity.
# The pins below only assert points that are farther than tol + the faceting error from the analytic surface, where
# inside / outside does not depend on that band rule; inside the band the device is compared with the oracle only.
def _solids(tmp_path):
    from tests.stl_util import torus_triangles, l_extrusion_triangles, in_l_extrusion
    rng = np.random.default_rng(0)
    out = []
    # sphere r = 0.7 around c: icosphere with 5120 faces, inscribed -> faceting error r * (1 - cos(half edge angle))
    c, r = np.array([0.1, -0.2, 0.3]), 0.7
    tri = icosphere_triangles(4, radius=r, center=c)
    pts = c + (rng.random((60000, 3)) - 0.5) * 2.2 * r
    d = np.linalg.norm(pts - c, axis=1) - r
    facet = 0.004 * r
    out.append(("sphere", tri, pts, np.where(d < -facet, 1, np.where(d > 0, -1, 0)), d))
    # torus R = 1, r = 0.35 (48 x 24 quads): distance to the analytic torus
    tri = torus_triangles(96, 48, 1.0, 0.35)
    pts = (rng.random((60000, 3)) - 0.5) * np.array([3.0, 3.0, 1.0])
    d = np.sqrt((np.sqrt(pts[:, 0] ** 2 + pts[:, 1] ** 2) - 1.0) ** 2 + pts[:, 2] ** 2) - 0.35
    facet = 0.003
    out.append(("torus", tri, pts, np.where(d < -facet, 1, np.where(d > 0, -1, 0)), d))
    # L-shaped extrusion: exact planar faces, no faceting error; includes lattice-aligned points
    tri = l_extrusion_triangles(0.5)
    pts = np.concatenate([(rng.random((40000, 3)) - 0.25) * np.array([3.0, 3.0, 1.0]),
                          np.stack(np.meshgrid(np.linspace(-0.5, 2.5, 25), np.linspace(-0.5, 2.5, 25),
                                               np.linspace(-0.25, 0.75, 9), indexing="ij"), -1).reshape(-1, 3)])
    out.append(("L", tri, pts, None, None))
    return out


def test_stl_closed_form_solids_pins(cuda, tmp_path):
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    from sparsespatialsampling_b200.geometry.device import nodes_inside
    from tests.stl_util import in_l_extrusion
    for name, tri, pts, want, dist in _solids(tmp_path):
        p = tmp_path / f"{name}.stl"
        write_binary_stl(p, tri)
        g = GeometrySTL3D(name, False, str(p))
        tol = g._tolerance
        got = nodes_inside(g, pt.from_numpy(pts)).numpy()
        # device == CPU oracle on every point, including the tolerance band
        assert np.array_equal(got, orc.points_inside(g, pts)), name
        if name == "L":
            cls = in_l_extrusion(pts, 0.5, margin=tol * 1.001 + 1e-6)    # fp32 STL coordinates: 1e-6 slack
        else:
            cls = np.where(want == 1, np.where(dist < -tol - 0.01, 1, 0), np.where(dist > tol, -1, 0))
        assert (got[cls == 1]).all() and (~got[cls == -1]).all(), name
        assert (cls == 1).sum() > 2000 and (cls == -1).sum() > 2000
        # band rule: on the exact planar faces of the L, points within tol of a face (outside) count as inside
        if name == "L":
            band = np.stack([np.full(50, 2.0 + 0.5 * tol), np.linspace(0.1, 0.9, 50), np.full(50, 0.25)], 1)
            far = band + np.array([1.0 * tol, 0, 0])
            assert nodes_inside(g, pt.from_numpy(band)).numpy().all()
            assert not nodes_inside(g, pt.from_numpy(far)).numpy().any()


@pytest.mark.parametrize("n_sub,n_cells", [(3, 3000), (5, 20000)])
def test_stl_tiled_kernel_equals_per_thread_path(cuda, tmp_path, n_sub, n_cells):
    """csrc/stl.cuh (node per thread, triangle tiles through shared memory, box rejects) against the single-point test
    in_stl evaluated per cell thread (stl_geoms = 0): same flags for cells, explicit nodes and points, all modes."""
    import ctypes
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.geometry import GeometrySTL3D, CubeGeometry
    from sparsespatialsampling_b200.geometry.device import GeometryTable
    p = tmp_path / "s.stl"
    write_binary_stl(p, icosphere_triangles(n_sub, radius=0.31, center=(0.5, 0.45, 0.55)))
    geoms = [CubeGeometry("domain", True, [0, 0, 0], [1, 1, 1]), GeometrySTL3D("body", False, str(p))]
    dev = pt.device("cuda")
    tab = GeometryTable(geoms, dev)
    assert tab.stl_geoms == 2
    lib = _lib.load()
    rng = np.random.default_rng(n_sub)
    level = pt.from_numpy(rng.integers(3, 7, n_cells).astype(np.int32)).to(dev)
    center = pt.from_numpy(rng.random((n_cells, 3))).to(dev)
    for refine_mode in (0, 1):
        outs = []
        for stl_geoms, meta in ((tab.stl_geoms, tab.stl_meta), (0, None)):
            inv = pt.empty(n_cells, dtype=pt.uint8, device=dev)
            _lib.check(lib.s3_cells_mask(_lib.ptr(center), _lib.ptr(level), None, 0, n_cells, 3, 1.0, _lib.ptr(tab.hdr),
                                         _lib.ptr(tab.par), tab.n, -1, refine_mode, 0, _lib.ptr(inv), None, None,
                                         stl_geoms, meta, _lib.stream_ptr()))
            outs.append(inv.cpu())
        assert pt.equal(outs[0], outs[1]) and 0 < int(outs[0].sum()) < n_cells
    pts = pt.from_numpy(rng.random((50000, 3))).to(dev)
    flags = []
    for stl_geoms, meta in ((tab.stl_geoms, tab.stl_meta), (0, None)):
        ins = pt.empty(pts.size(0), dtype=pt.uint8, device=dev)
        _lib.check(lib.s3_points_inside(_lib.ptr(pts), pts.size(0), 3, _lib.ptr(tab.hdr), _lib.ptr(tab.par), 1,
                                        _lib.ptr(ins), stl_geoms, meta, _lib.stream_ptr()))
        flags.append(ins.cpu())
    assert pt.equal(flags[0], flags[1])
    vol = float(flags[0].float().mean())
    assert abs(vol - 4.0 / 3.0 * np.pi * 0.31 ** 3) < 0.01                   # Monte-Carlo volume of the sphere


@pytest.mark.parametrize("flip_every", [0, 3])
def test_stl_rays_through_edges_and_vertices(cuda, tmp_path, flip_every):
    """Rays that hit shared edges, face diagonals and mesh vertices EXACTLY (after the test's own nudge): device (tiled
    kernel and per-thread path) == oracle == closed form, independent of the triangle winding."""
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.geometry import GeometrySTL3D
    from sparsespatialsampling_b200.geometry.device import GeometryTable, nodes_inside
    from tests.stl_util import two_box_exact_hit_case
    tri, pts, want = two_box_exact_hit_case(flip_every)
    p = tmp_path / "two.stl"
    write_binary_stl(p, tri)
    g = GeometrySTL3D("two", False, str(p))
    got = nodes_inside(g, pt.from_numpy(pts)).numpy()
    assert np.array_equal(got, want) and np.array_equal(got, orc.points_inside(g, pts))
    dev = pt.device("cuda")
    tab = GeometryTable([g], dev)
    lib = _lib.load()
    d_pts = pt.from_numpy(pts).to(dev)
    for stl_geoms, meta in ((tab.stl_geoms, tab.stl_meta), (0, None)):
        ins = pt.empty(d_pts.size(0), dtype=pt.uint8, device=dev)
        _lib.check(lib.s3_points_inside(_lib.ptr(d_pts), d_pts.size(0), 3, _lib.ptr(tab.hdr), _lib.ptr(tab.par), 0,
                                        _lib.ptr(ins), stl_geoms, meta, _lib.stream_ptr()))
        assert np.array_equal(ins.cpu().numpy().astype(bool), want)
