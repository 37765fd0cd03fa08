"""
Generates the golden fixtures under tests/golden/ by running the REFERENCE implementation
(/root/reference, JanisGeise/sparseSpatialSampling v1.0.0) in the build container, and checks the CPU oracle
(oracle/s3_oracle.py) against it on the way. Run once by hand:  python tests/golden/make_golden.py

The reference needs flowtorch / shapely / pyvista / pymeshfix / h5py at import time; they are not installed here, so
stand-in modules are written to a temp dir (flowtorch.data.mask_box = inclusive box, mask_sphere = ||v-c|| <= r; the
others are empty shells that are never called by the cases below). Neither /root/reference nor this script is needed
at test time -- the tests only read the .npz files written here.
"""
import os
import sys
import tempfile
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def write_stubs() -> str:
    d = tempfile.mkdtemp(prefix="s3_stubs_")
    files = {
        "flowtorch/__init__.py": "",
        "flowtorch/data.py": """
            import torch as pt
            def mask_box(vertices, lower, upper):
                mask = pt.ones(vertices.shape[0], dtype=pt.bool)
                for i, (lo, up) in enumerate(zip(lower, upper)):
                    mask = mask & (vertices[:, i] >= lo) & (vertices[:, i] <= up)
                return mask
            def mask_sphere(vertices, center, radius):
                loc = pt.tensor(center, dtype=vertices.dtype)
                return (vertices - loc).norm(dim=1) <= radius
            class FOAMDataloader: pass
            """,
        "flowtorch/analysis.py": "class SVD: pass\n",
        "shapely/__init__.py": "class Point: pass\nclass Polygon: pass\n",
        "pyvista/__init__.py": "class PolyData: pass\ndef read(*a, **k): raise NotImplementedError\n",
        "pymeshfix/__init__.py": "class MeshFix: pass\n",
        "h5py/__init__.py": "class File: pass\n",
    }
    for name, body in files.items():
        p = os.path.join(d, name)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(textwrap.dedent(body))
    return d


def case_definitions(geo):
    """name -> dict(coords, metric, geometries(factory taking the geometry module), kwargs)"""
    import torch as pt
    import synth
    cases = {}

    x = synth.cylinder2d_cloud(4000, seed=1)
    cases["g2d_metric"] = dict(
        coords=x, metric=synth.wake_metric(x),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
                         g.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True)],
        kwargs=dict(uniform_level=4, min_metric=0.6, n_cells_iter_start=20))

    x = synth.airfoil2d_cloud(6000, seed=2)
    cases["g2d_ncells"] = dict(
        coords=x, metric=synth.wake_metric(x, xc=1.0, yc=0.0),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.AIRFOIL2D["lower"], synth.AIRFOIL2D["upper"]),
                         g.TriangleGeometry("wedge", False, [[0.0, 0.0], [1.0, 0.06], [1.0, -0.06]], refine=True,
                                            min_refinement_level=8)],
        kwargs=dict(uniform_level=4, n_cells=1500, n_cells_iter_start=40))

    x = synth.cylinder3d_cloud(6000, seed=3)
    zmax = synth.CYL3D["upper"][2]
    cases["g3d_metric"] = dict(
        coords=x, metric=synth.wake_metric(x, xc=0.8, yc=1.0),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL3D["lower"], synth.CYL3D["upper"]),
                         g.CylinderGeometry3D("cylinder", False, [[0.8, 1.0, 0.0], [0.8, 1.0, zmax]], 0.25,
                                              refine=True),
                         g.PrismGeometry3D("prism", False, [[[1.6, 0.4, 0.0], [2.2, 0.4, 0.0], [1.9, 1.0, 0.0]],
                                                            [[1.6, 0.4, zmax], [2.2, 0.4, zmax], [1.9, 1.0, zmax]]]),
                         g.TetrahedronGeometry3D("tet", False, [[0.0, 0.0, 0.0], [0.7, 0.0, 0.0], [0.0, 0.7, 0.0],
                                                                [0.0, 0.0, 0.7]]),
                         g.PyramidGeometry3D("pyr", False, [[1.2, 1.4, 0.0], [2.0, 1.4, 0.0], [2.0, 2.0, 0.0],
                                                            [1.2, 2.0, 0.0], [1.6, 1.7, 0.9]])],
        kwargs=dict(uniform_level=3, min_metric=0.5, n_cells_iter_start=10))
    # delta-level constraint in the metric-based loop (s_cube.py:447-506, 611-618). No geometry refinement here: in
    # that phase the reference follows neighbour pointers that can be stale (a child inherits the wrong sibling of a
    # neighbour refined after the pointer was set), which the geometric restatement does not reproduce -- DESIGN.md 6.
    x = synth.cylinder2d_cloud(4000, seed=8)
    cases["g2d_delta"] = dict(
        coords=x, metric=synth.wake_metric(x),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
                         g.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"])],
        kwargs=dict(uniform_level=3, min_metric=0.6, n_cells_iter_start=10, max_delta_level=True))
    # delta-level constraint where the reference's neighbour pointers go stale (a child inherits the same-position child of
    # a neighbour pointer that still refers to a coarser cell refined later): 3-D metric loop, and 2-D with geometry
    # refinement. Only the pointer replay (oracle/topology_oracle.py, csrc/topology.cu) reproduces these.
    x = synth.cylinder3d_cloud(6000, seed=3)
    cases["g3d_delta"] = dict(
        coords=x, metric=synth.wake_metric(x, xc=0.8, yc=1.0),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL3D["lower"], synth.CYL3D["upper"]),
                         g.CylinderGeometry3D("cylinder", False, [[0.8, 1.0, 0.0], [0.8, 1.0, zmax]], 0.25)],
        kwargs=dict(uniform_level=2, min_metric=0.55, n_cells_iter_start=10, max_delta_level=True))
    x = synth.cylinder2d_cloud(4000, seed=12)
    cases["g2d_delta_geo"] = dict(
        coords=x, metric=synth.wake_metric(x),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"], refine=True,
                                        min_refinement_level=6),
                         g.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True,
                                          min_refinement_level=8)],
        kwargs=dict(uniform_level=3, min_metric=0.5, n_cells_iter_start=10, max_delta_level=True))
    # the second stopping rule (s_cube.py:263-284: metric within reach_at_least of the target and improving by less than
    # relTol), a cells-per-iteration schedule that decays (n_cells_iter_end < start, s_cube.py:286-315) and a box used as a
    # body (keep_inside=False). pre_select=True cannot be pinned with analytic shapes: no cell is ever removed then, every
    # node id stays in use and the reference fails in renumber_node_indices_parallel (numba cannot type the empty set).
    x = synth.cylinder2d_cloud(5000, seed=21)
    cases["g2d_reltol"] = dict(
        coords=x, metric=synth.wake_metric(x),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
                         g.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"]),
                         g.CubeGeometry("block", False, [1.2, 0.05], [1.5, 0.15])],
        kwargs=dict(uniform_level=3, min_metric=0.9, n_cells_iter_start=40, n_cells_iter_end=4, relTol=4e-3,
                    reach_at_least=0.5))
    # 3-D: delta-level constraint together with geometry refinement of a sphere and of the domain box, stop by n_cells
    x = synth.cylinder3d_cloud(5000, seed=22)
    cases["g3d_delta_geo"] = dict(
        coords=x, metric=synth.wake_metric(x, xc=0.8, yc=1.0),
        geoms=lambda g: [g.CubeGeometry("domain", True, synth.CYL3D["lower"], synth.CYL3D["upper"]),
                         g.SphereGeometry("ball", False, [0.8, 1.0, 0.15], 0.12, refine=True, min_refinement_level=5)],
        kwargs=dict(uniform_level=2, n_cells=700, n_cells_iter_start=12, max_delta_level=True))
    return cases


def run_reference_case(name, case, geo_ref):
    import torch as pt
    from sparseSpatialSampling.s_cube import SamplingTree
    from sparseSpatialSampling.export import interpolate_data
    from sklearn.neighbors import NearestNeighbors
    geoms = case["geoms"](geo_ref)
    tree = SamplingTree(case["coords"], case["metric"], geoms, n_jobs=4, **case["kwargs"])
    gain0 = float(np.asarray(tree._cells[0].gain).reshape(-1)[0])
    tree.refine()
    leaf_list = list(tree._leaf_cells)
    assert leaf_list == sorted(leaf_list), "reference leaf-set order is not ascending in this case"
    assert leaf_list == [c.index for c in tree._cells if c.leaf_cell()]
    info = tree.data_final_mesh
    out = dict(
        centers=tree.all_centers.numpy(), levels=tree.all_levels.numpy(), faces=tree.face_ids.numpy(),
        vertices=tree.all_nodes.numpy(), leaf_index=np.asarray(leaf_list, dtype=np.int64),
        cells_per_iter=np.asarray(info["cells_per_iter"], dtype=np.int64),
        metric_per_iter=np.asarray(info["metric_per_iter"], dtype=np.float64),
        iterations=np.int64(info["iterations"]), n_cells=np.int64(info["n_cells"]),
        min_level=np.int64(info["min_level"]), max_level=np.int64(info["max_level"]),
        width=np.float64(info["size_initial_cell"]), gain0=np.float64(gain0),
        leaf_gain=np.asarray([float(tree._cells[i].gain) for i in leaf_list]),
        leaf_metric=np.asarray([float(tree._cells[i].metric) for i in leaf_list]),
        n_cells_total=np.int64(len(tree._cells)),
    )
    # export stage on the generated grid (export.py:403-468)
    import synth
    d = case["coords"].shape[1]
    k = 8 if d == 2 else 26
    nn = NearestNeighbors(n_neighbors=k).fit(case["coords"].numpy())
    dist, idx = nn.kneighbors(tree.all_centers.numpy())
    w = 1.0 / pt.clamp(pt.from_numpy(dist), min=1e-12)
    w /= w.sum(axis=1, keepdim=True)
    field = synth.wake_field(case["coords"], 0, 12, 12, components=2)
    interp = interpolate_data(w, pt.from_numpy(idx), field, 100000)
    out.update(knn_idx=idx.astype(np.int32), knn_w=w.numpy(), interp=interp.numpy())
    return tree, out


def reference_neighbour_table(d: int):
    """
    The table behind the reference's hand-written ``_assign_neighbors`` (s_cube.py:904-1186), read off by running it on a
    parent whose neighbours and their children are labelled dummies: table[c][s] = ("sibling", j) | ("poc", P, j).
    """
    from sparseSpatialSampling.s_cube import SamplingTree, Cell
    nnb, nch = (8, 4) if d == 2 else (26, 8)

    class _Self:
        _n_dimensions = d

    def make_parent(with_children: bool):
        nbs = []
        for P in range(nnb):
            n = Cell(100 + P, None, nnb * [None], None, 1, dimensions=d)
            if with_children:
                n.children = tuple(Cell(1000 + 10 * P + j, n, nnb * [None], None, 2, dimensions=d) for j in range(nch))
            nbs.append(n)
        parent = Cell(0, None, nbs, None, 1, dimensions=d)
        kids = [Cell(1 + j, parent, nnb * [None], None, 2, dimensions=d) for j in range(nch)]
        return parent, kids

    p1, k1 = make_parent(True)
    SamplingTree._assign_neighbors(_Self(), p1, children=k1)
    p0, k0 = make_parent(False)
    SamplingTree._assign_neighbors(_Self(), p0, children=k0)
    table = []
    for c in range(nch):
        row = []
        for s_ in range(nnb):
            a, b = k1[c].nb[s_], k0[c].nb[s_]
            if 1 <= a.index <= nch:
                assert b.index == a.index
                row.append(("sibling", a.index - 1))
            else:
                P, j = divmod(a.index - 1000, 10)
                assert b.index == 100 + P, "parent_or_child must fall back to the same parent neighbour"
                row.append(("poc", P, j))
        table.append(row)
    return table


def check_oracle_against_reference(name, case, geo_ref, ref_out):
    from oracle import s3_oracle as orc
    from oracle.topology_oracle import OracleTopology
    geoms = case["geoms"](geo_ref)
    kw = dict(case["kwargs"])
    tree = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), geoms, **kw, sdm_order=1,
                          topology=OracleTopology).refine()
    assert np.array_equal(tree.face_ids, ref_out["faces"]), f"{name}: faces (node ids) differ"
    assert np.array_equal(tree.all_nodes, ref_out["vertices"]), f"{name}: vertices differ"
    assert tree.leaf_order == ref_out["leaf_index"].tolist(), f"{name}: leaf numbering differs"
    assert np.array_equal(tree.all_centers, ref_out["centers"]), f"{name}: centres differ"
    assert np.array_equal(tree.all_levels, ref_out["levels"]), f"{name}: levels differ"
    assert tree.n_cells_log == ref_out["cells_per_iter"].tolist(), f"{name}: cells_per_iter differs"
    assert tree.gain0 == float(ref_out["gain0"]), f"{name}: gain0 differs"
    g = np.asarray([tree.gain[i] for i in tree.leaf_order])
    m = np.asarray([tree.metric[i] for i in tree.leaf_order])
    assert np.array_equal(g, ref_out["leaf_gain"]), f"{name}: leaf gains differ"
    assert np.array_equal(m, ref_out["leaf_metric"]), f"{name}: leaf metrics differ"
    np.testing.assert_allclose(tree.metric_log, ref_out["metric_per_iter"], rtol=1e-12)
    # export stage
    d, i = orc.knn_search(case["coords"].numpy(), ref_out["centers"], ref_out["knn_idx"].shape[1])
    assert np.array_equal(i, ref_out["knn_idx"].astype(np.int64)), f"{name}: export KNN indices differ"
    np.testing.assert_allclose(orc.export_weights(d), ref_out["knn_w"], rtol=1e-14)
    import synth
    field = synth.wake_field(case["coords"], 0, 12, 12, components=2).numpy()
    np.testing.assert_allclose(orc.interpolate(ref_out["knn_w"], i, field), ref_out["interp"], rtol=1e-13, atol=1e-15)
    print(f"   oracle == reference for {name}: {len(tree.leaf_order)} leaves, {tree.iterations} iterations, "
          f"{len(tree.center)} cells created")
    return tree


def geometry_pins(geo_ref):
    """check_cell of every analytic reference class on random cells around the shape (both modes) + oracle check."""
    import torch as pt
    from oracle import s3_oracle as orc
    rng = np.random.default_rng(42)
    shapes = {
        "cube2d": (lambda g, ki: g.CubeGeometry("c", ki, [0.0, 0.0], [1.0, 1.0]), 2),
        "cube3d": (lambda g, ki: g.CubeGeometry("c", ki, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]), 3),
        "sphere2d": (lambda g, ki: g.SphereGeometry("s", ki, [0.5, 0.5], 0.45), 2),
        "sphere3d": (lambda g, ki: g.SphereGeometry("s", ki, [0.5, 0.5, 0.5], 0.45), 3),
        "cylinder": (lambda g, ki: g.CylinderGeometry3D("cy", ki, [[0.1, 0.2, 0.3], [0.9, 0.7, 0.6]], 0.3), 3),
        "cone": (lambda g, ki: g.CylinderGeometry3D("co", ki, [[0.5, 0.5, 0.0], [0.5, 0.5, 1.0]], [0.4, 0.1]), 3),
        "triangle": (lambda g, ki: g.TriangleGeometry("t", ki, [[0.0, 0.0], [1.0, 0.1], [0.4, 0.9]]), 2),
        "prism": (lambda g, ki: g.PrismGeometry3D("p", ki, [[[0.0, 0.0, 0.1], [1.0, 0.1, 0.1], [0.4, 0.9, 0.1]],
                                                            [[0.0, 0.0, 0.8], [1.0, 0.1, 0.8], [0.4, 0.9, 0.8]]]), 3),
        "tetra": (lambda g, ki: g.TetrahedronGeometry3D("te", ki, [[0.0, 0.0, 0.0], [1.0, 0.1, 0.0], [0.3, 0.9, 0.1],
                                                                  [0.4, 0.3, 0.95]]), 3),
        "pyramid": (lambda g, ki: g.PyramidGeometry3D("py", ki, [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0],
                                                                [0.0, 1.0, 0.0], [0.5, 0.5, 1.0]]), 3),
    }
    out = {}
    for name, (factory, dim) in shapes.items():
        n = 400
        dirs = orc.DIRS_2D if dim == 2 else orc.DIRS_3D
        centers = rng.random((n, dim)) * 1.6 - 0.3
        half = 2.0 ** -rng.integers(2, 7, n)
        nodes = centers[:, None, :] + dirs[None, :, :] * half[:, None, None]
        # lattice-like cells whose nodes land exactly on nice numbers (boundary-inclusive behaviour)
        nodes[:40] = np.round(nodes[:40] * 8) / 8
        res = np.zeros((n, 4), dtype=bool)
        for ci, (ki, rf) in enumerate([(True, False), (False, False), (True, True), (False, True)]):
            g = factory(geo_ref, ki)
            for t in range(n):
                r = g.check_cell(pt.from_numpy(nodes[t]), rf)
                res[t, ci] = r
                assert orc.check_cell(g, nodes[t], rf) == r, f"oracle mask differs from reference: {name} cell {t}"
        out[f"{name}_nodes"] = nodes
        out[f"{name}_invalid"] = res
        print(f"   oracle == reference for geometry {name}: {res.sum(0).tolist()} invalid of {n}")
    return out


def main():
    stubs = write_stubs()
    sys.path[:0] = [stubs, REF, ROOT]
    os.environ["PYTHONPATH"] = os.pathsep.join([stubs, REF, ROOT, os.environ.get("PYTHONPATH", "")])
    import sparseSpatialSampling.geometry as geo_ref
    from oracle.topology_oracle import neighbour_table
    for d in (2, 3):
        assert reference_neighbour_table(d) == neighbour_table(d), f"neighbour table differs from the reference ({d}-D)"
    print("   oracle neighbour table == the reference's _assign_neighbors (4*8 and 8*26 entries)")
    if not [a for a in sys.argv[1:] if not a.startswith("-")]:
        np.savez_compressed(os.path.join(HERE, "geometry_pins.npz"), **geometry_pins(geo_ref))
    cases = case_definitions(geo_ref)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, case in cases.items():
        if only and name not in only:
            continue
        print(f"== reference run: {name} ({case['coords'].shape[0]} points)")
        _, out = run_reference_case(name, case, geo_ref)
        check_oracle_against_reference(name, case, geo_ref, out)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(f"   wrote {name}.npz: n_cells={int(out['n_cells'])} iterations={int(out['iterations'])}")


if __name__ == "__main__":
    main()
