"""CPU tests of the host-side logic: geometry set-up, parameter packing, oracle vs reference golden fixtures."""
import os

import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc
from tests.golden.make_golden import case_definitions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["g2d_metric", "g2d_ncells", "g3d_metric", "g2d_delta", "g2d_reltol"])
def test_oracle_reproduces_reference_golden(name):
    # pins the oracle to outputs of the reference itself (leaf cells, numbering, gains, metrics: bit-exact)
    import sparsespatialsampling_b200.geometry as geo
    case = case_definitions(geo)[name]
    ref = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    tree = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"],
                          sdm_order=1).refine()
    assert tree.leaf_order == ref["leaf_index"].tolist()
    assert np.array_equal(tree.all_centers, ref["centers"])
    assert np.array_equal(tree.all_levels, ref["levels"])
    assert tree.n_cells_log == ref["cells_per_iter"].tolist()
    assert tree.gain0 == float(ref["gain0"])
    assert np.array_equal(np.asarray([tree.gain[i] for i in tree.leaf_order]), ref["leaf_gain"])
    assert np.array_equal(np.asarray([tree.metric[i] for i in tree.leaf_order]), ref["leaf_metric"])
    np.testing.assert_allclose(tree.metric_log, ref["metric_per_iter"], rtol=1e-12)
    d, i = orc.knn_search(case["coords"].numpy(), ref["centers"], ref["knn_idx"].shape[1])
    assert np.array_equal(i, ref["knn_idx"].astype(np.int64))


def test_oracle_masks_reproduce_reference_golden():
    from tests.test_geometry_gpu import _shapes
    pins = np.load(os.path.join(GOLDEN, "geometry_pins.npz"))
    for name, factory in _shapes().items():
        nodes, expected = pins[f"{name}_nodes"], pins[f"{name}_invalid"]
        for ci, (ki, rf) in enumerate([(True, False), (False, False), (True, True), (False, True)]):
            g = factory(ki)
            got = np.array([orc.check_cell(g, nodes[t], rf) for t in range(nodes.shape[0])])
            assert np.array_equal(got, expected[:, ci]), (name, ki, rf)


def test_device_parameter_blocks_match_oracle_packing():
    # the product's device_params() and the oracle's attribute-based packing describe the same shape
    from tests.test_geometry_gpu import _shapes
    for name, factory in _shapes().items():
        g = factory(False)
        t_id, par, n_extra = g.device_params()
        o_id, o_par, o_extra = orc.geometry_params(g)
        assert (t_id, n_extra) == (o_id, o_extra), name
        assert par == o_par, name


def test_geometry_properties_follow_reference_conventions():
    import sparsespatialsampling_b200.geometry as g
    dom = g.CubeGeometry("domain", True, [0, 0], [2.2, 0.41])
    assert dom.main_width == 2.2 and dom.type == "cube" and dom.keep_inside is True
    assert pt.equal(dom.center, pt.tensor([1.1, 0.205], dtype=pt.float64))
    cyl = g.CylinderGeometry3D("c", False, [[0.8, 1.0, 0.0], [0.8, 1.0, 0.3]], 0.05)
    # end points are rounded to float32 like the reference (cylinder_geometry.py:52)
    assert cyl._position.dtype == pt.float32 and cyl._axis.dtype == pt.float64
    assert abs(cyl.main_width - 0.3) < 1e-6
    prism = g.PrismGeometry3D("p", False, [[[0.0, 0.0, 0.1], [1.0, 0.1, 0.1], [0.4, 0.9, 0.1]],
                                           [[0.0, 0.0, 0.8], [1.0, 0.1, 0.8], [0.4, 0.9, 0.8]]])
    assert prism._dim.tolist() == [0, 1] and abs(prism.center[2].item() - 0.45) < 1e-12
    pyr = g.PyramidGeometry3D("py", False, [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0], [0.0, 1.0, 0.0],
                                            [0.5, 0.5, 1.0]])
    assert pyr._apex_idx == 4 and len(pyr._tets) == 2
    tri = g.TriangleGeometry("t", False, [[0.0, 0.0], [1.0, 0.06], [1.0, -0.06]])
    assert tri._points[1].dtype == pt.float64 and tri._points[1][1].item() == 0.06     # parsed as fp64, not fp32


def test_sum_order_probe_matches_this_host():
    from sparsespatialsampling_b200.s_cube import probe_sum_order_8
    assert probe_sum_order_8() in (0, 1)


def test_snapshot_windows_partition_the_time_axis():
    from sparsespatialsampling_b200.parallel import snapshot_window
    for n, w in [(1000, 8), (1001, 8), (7, 8), (2000, 3), (1, 1)]:
        wins = [snapshot_window(n, w, r) for r in range(w)]
        assert wins[0][0] == 0 and wins[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(wins, wins[1:]))
        sizes = [b - a for a, b in wins]
        assert max(sizes) - min(sizes) <= 1


def test_surface_geometry_oracle_meets_reference_unit_tests(tmp_path):
    # the reference's own expectations for GeometrySTL3D / GeometryCoordinates2D (tests/test_geometry_STL.py:31-63,
    # tests/test_coordinates_2d_geometry.py:37-71), evaluated with the CPU oracle of the documented restatement
    import sparsespatialsampling_b200.geometry as g
    from tests.stl_util import write_binary_stl, cube_triangles
    from tests.test_geometry_gpu import DummyCells
    p = tmp_path / "cube.stl"
    write_binary_stl(p, cube_triangles())
    c = DummyCells()
    for keep_inside, expected in [(False, [False, True, False]), (True, [True, False, False])]:
        stl = g.GeometrySTL3D("cube", keep_inside, str(p))
        got = [orc.check_cell(stl, getattr(c, n).double().numpy())
               for n in ("cell_outside_3D", "cell_inside_3D", "cell_partially_3D")]
        assert got == expected
        poly = g.GeometryCoordinates2D("square", keep_inside, [(-1, -1), (-1, 1.25), (1.25, 1.25), (1.25, -1)])
        got = [orc.check_cell(poly, getattr(c, n).double().numpy())
               for n in ("cell_outside_2D", "cell_inside_2D", "cell_partially_2D")]
        assert got == expected
    assert g.GeometrySTL3D("cube", False, str(p)).device_params()[2] == 12


def test_top_eigenpairs_subspace_iteration_and_fallback():
    """svd.top_eigenpairs (pure torch, runs on the CPU as well): exact leading eigenpairs for decaying spectra, early
    bail-out (None -> full eigh) for a flat tail inside the requested pairs and when r is not small against T."""
    from sparsespatialsampling_b200.svd import top_eigenpairs
    pt.manual_seed(0)
    t = 600
    u, _ = pt.linalg.qr(pt.randn(t, t, dtype=pt.float64))
    decaying = (1.0 / pt.arange(1, t + 1, dtype=pt.float64)) ** 3
    g = (u * decaying) @ u.T
    for r in (4, 30):
        w, x = top_eigenpairs(g, r)
        ref_w, ref_v = pt.linalg.eigh(g)
        ref_w, ref_v = ref_w.flip(0)[:r], ref_v.flip(1)[:, :r]
        assert float(((w - ref_w).abs() / ref_w).max()) < 1e-10
        assert float((x.T @ x - pt.eye(r, dtype=pt.float64)).abs().max()) < 1e-10
        assert float((x * ref_v).sum(0).abs().min()) > 1 - 1e-8          # same vectors up to sign
    flat = pt.cat([pt.logspace(0, -3, 10, dtype=pt.float64), 1e-4 * pt.logspace(0, -2, t - 10, dtype=pt.float64)])
    assert top_eigenpairs((u * flat) @ u.T, 30) is None                   # slowly decaying tail: not worth iterating
    assert top_eigenpairs(g, 200) is None                                 # r not small against T


@pytest.mark.parametrize("flip_every", [0, 3])
def test_stl_oracle_rays_through_edges_and_vertices(tmp_path, flip_every):
    """The restated STL inside test counts a ray that runs exactly through a shared edge / a vertex of the mesh once
    per surface crossing (half-open edge rule), whatever the winding of the triangles."""
    import sparsespatialsampling_b200.geometry as g
    from tests.stl_util import write_binary_stl, two_box_exact_hit_case
    tri, pts, want = two_box_exact_hit_case(flip_every)
    p = tmp_path / "two.stl"
    write_binary_stl(p, tri)
    stl = g.GeometrySTL3D("two", False, str(p))
    assert np.array_equal(orc.points_inside(stl, pts), want)
