"""
The reference's 2-D cylinder example (examples/s3_for_cylinder2D_Re100.py:42-73) end to end through the public classes,
on synthetic data: grid generation -> ExportData (p in one call, U in two snapshot batches, one HDF5/XDMF pair) ->
write_svd_s_cube_to_file -> Dataloader reads fields and modes back. Every stage is cross-checked against the CPU oracle.
"""
import os

import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

pytestmark = pytest.mark.gpu


def test_cylinder2d_example_flow(cuda, tmp_path):
    import synth
    import sparsespatialsampling_b200 as s3
    from sparsespatialsampling_b200 import (SparseSpatialSampling, ExportData, Dataloader, write_svd_s_cube_to_file)
    from sparsespatialsampling_b200.geometry import CubeGeometry, SphereGeometry
    n_t = 24
    coord = synth.cylinder2d_cloud(6000, seed=31)
    u = synth.wake_field(coord, 0, n_t, n_t, components=2)                 # [N, 2, T] fp32
    p = synth.wake_field(coord, 0, n_t, n_t, components=1)                 # [N, 1, T]
    write_times = [f"{0.1 * i:.1f}" for i in range(n_t)]
    domain = CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"])
    cylinder = SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True,
                              min_refinement_level=7)
    metric = pt.mean(u.abs().sum(1), 1).to(pt.float64)
    s_cube = SparseSpatialSampling(coord, metric, [domain, cylinder], str(tmp_path), "metric_0.60", "cylinder2D",
                                   uniform_levels=4, min_metric=0.6, n_jobs=4)
    s_cube.execute_grid_generation()
    nc = s_cube.centers.size(0)
    assert nc > 200 and s_cube.faces.shape == (nc, 4)

    # grid: same leaves as the oracle's restatement of the reference loop
    tree = orc.OracleTree(coord.numpy(), metric.numpy(), [domain, cylinder], uniform_level=4, min_metric=0.6,
                          sdm_order=1).refine()
    assert np.array_equal(tree.all_centers, s_cube.centers.numpy())
    assert np.array_equal(tree.all_levels, s_cube.levels.numpy())

    export = ExportData(s_cube, write_new_file_for_each_field=False, write_times=write_times)
    export.export(coord, p, "p")
    export.export(coord, u[:, :, :10], "U", n_snapshots_total=n_t)        # two batches of snapshots
    export.export(coord, u[:, :, 10:], "U", n_snapshots_total=n_t)
    export.synchronize() if hasattr(export, "synchronize") else None

    loader = Dataloader(str(tmp_path), "metric_0.60.h5")
    assert loader.write_times == write_times
    assert sorted(loader.field_names[write_times[0]]) == ["U", "p"]
    assert np.array_equal(loader.vertices.numpy(), s_cube.centers.numpy())
    p_s, u_s = loader.load_snapshot("p"), loader.load_snapshot("U")
    assert tuple(p_s.shape) == (nc, n_t) and tuple(u_s.shape) == (nc, 2, n_t)
    # interpolation against the oracle (export.py:403-468)
    d, idx = orc.knn_search(coord.numpy(), s_cube.centers.numpy(), 8)
    w = orc.export_weights(d)
    ref_u = orc.interpolate(w, idx, u.numpy())
    ref_p = orc.interpolate(w, idx, p.numpy())[:, 0]
    assert np.abs(u_s.numpy() - ref_u).max() <= 1e-5 * np.abs(u.numpy()).max()
    assert np.abs(p_s.numpy() - ref_p).max() <= 1e-5 * np.abs(p.numpy()).max()
    assert os.path.exists(tmp_path / "metric_0.60.xdmf")

    # SVD of both fields from the written file (utils.py:349-413)
    write_svd_s_cube_to_file(["p", "U"], str(tmp_path), "metric_0.60", export.new_file, 5, rank=int(1e5),
                             t_start=0.4)
    keep = [i for i, t in enumerate(write_times) if float(t) >= 0.4]
    for name, ref_field in (("p", ref_p[:, keep]), ("U", ref_u[:, :, keep])):
        out = Dataloader(str(tmp_path), f"metric_0.60_{name}_svd.h5")
        st = out._store()
        s = np.asarray(st.read("constant/s")).reshape(-1)
        assert s.shape[0] == len(keep) and np.all(np.diff(s) <= 1e-6 * s[0])
        s_ref, u_ref, v_ref = orc.compute_svd(ref_field.astype(np.float32), loader.weights.numpy(), len(keep))
        assert np.abs(s[:5] - s_ref[:5]).max() <= 1e-4 * s_ref[0]
        mode1 = np.asarray(st.read("constant/mode_1")).reshape(-1)
        ref1 = np.asarray(u_ref)[..., 0].reshape(-1)
        assert abs(mode1 @ ref1) / (np.linalg.norm(mode1) * np.linalg.norm(ref1)) >= 1 - 1e-4
        assert os.path.exists(tmp_path / f"metric_0.60_{name}_svd.xdmf")


def test_cylinder3d_example_flow_snapshot_wise(cuda, tmp_path):
    # examples/s3_for_cylinder3D_Re3900.py:106-141: 3-D cloud, CylinderGeometry3D longer than the domain, fields exported
    # snapshot by snapshot (batch size 1, utils.py:210-226), values at cell centres AND vertices, SVD of the scalar field
    import synth
    from sparsespatialsampling_b200 import (SparseSpatialSampling, ExportData, Dataloader, write_svd_s_cube_to_file)
    from sparsespatialsampling_b200.geometry import CubeGeometry, CylinderGeometry3D
    n_t = 6
    coord = synth.cylinder3d_cloud(5000, seed=41)
    p = synth.wake_field(coord, 0, n_t, n_t, components=1, xc=0.8, yc=1.0)
    metric = synth.wake_metric(coord, xc=0.8, yc=1.0)
    write_times = [str(i) for i in range(n_t)]
    geometry = [CubeGeometry("domain", True, synth.CYL3D["lower"], synth.CYL3D["upper"]),
                CylinderGeometry3D("cylinder", False, [(0.8, 1.0, -1), (0.8, 1.0, 1)], 0.25, refine=True)]
    s_cube = SparseSpatialSampling(coord, metric, geometry, str(tmp_path), "c3d_metric_0.50", "cylinder",
                                   uniform_levels=3, min_metric=0.5, n_jobs=8, max_delta_level=False)
    s_cube.execute_grid_generation()
    nc, nv = s_cube.centers.size(0), s_cube.vertices.size(0)
    assert s_cube.faces.shape == (nc, 8) and int(s_cube.faces.max()) == nv - 1
    tree = orc.OracleTree(coord.numpy(), metric.numpy(), geometry, uniform_level=3, min_metric=0.5, sdm_order=1).refine()
    assert np.array_equal(tree.all_centers, s_cube.centers.numpy())

    export = ExportData(s_cube, interpolate_at_vertices=True, write_times=write_times)
    for i in range(n_t):                                                  # batch size 1
        export.export(coord, p[:, :, i:i + 1], "p", n_snapshots_total=n_t)
    loader = Dataloader(str(tmp_path), "c3d_metric_0.50.h5")
    assert loader.write_times == write_times
    p_s = loader.load_snapshot("p")
    assert tuple(p_s.shape) == (nc, n_t)
    d, idx = orc.knn_search(coord.numpy(), s_cube.centers.numpy(), 26)
    ref_p = orc.interpolate(orc.export_weights(d), idx, p.numpy())[:, 0]
    assert np.abs(p_s.numpy() - ref_p).max() <= 1e-5 * np.abs(p.numpy()).max()
    # the same field at the vertices of the grid (export.py:226-228, 437-444)
    st = loader._store()
    at_vertices = np.stack([np.asarray(st.read(f"data/{t}/p_vertices")).reshape(-1) for t in write_times], axis=1)
    dv, iv = orc.knn_search(coord.numpy(), s_cube.vertices.numpy(), 26)
    ref_v = orc.interpolate(orc.export_weights(dv), iv, p.numpy())[:, 0]
    assert at_vertices.shape == (nv, n_t)
    assert np.abs(at_vertices - ref_v).max() <= 1e-5 * np.abs(p.numpy()).max()

    write_svd_s_cube_to_file(["p"], str(tmp_path), "c3d_metric_0.50", export.new_file, 3, rank=int(1e5))
    out = Dataloader(str(tmp_path), "c3d_metric_0.50_p_svd.h5")._store()
    s = np.asarray(out.read("constant/s")).reshape(-1)
    s_ref, u_ref, _ = orc.compute_svd(ref_p.astype(np.float32), loader.weights.numpy(), n_t)
    assert s.shape[0] == n_t and np.abs(s[:3] - s_ref[:3]).max() <= 1e-4 * s_ref[0]
    assert sorted(k for k in out.keys("constant") if k.startswith("mode_")) == ["mode_1", "mode_2", "mode_3"]


def test_oat15_example_flow_polygon_bodies(cuda, tmp_path):
    # examples/s3_for_OAT15_airfoil.py:100-133: two bodies given as closed coordinate rings (GeometryCoordinates2D),
    # write times passed as floats, a scalar field [N, T] unsqueezed to [N, 1, T], SVD called with the field name as str
    import synth
    from sparsespatialsampling_b200 import (SparseSpatialSampling, ExportData, Dataloader, write_svd_s_cube_to_file)
    from sparsespatialsampling_b200.geometry import CubeGeometry, GeometryCoordinates2D
    n_t = 12
    xz = synth.airfoil2d_cloud(6000, seed=51)
    th = np.linspace(0.0, 2.0 * np.pi, 41)
    front = np.stack([0.5 + 0.5 * np.cos(th), 0.05 * np.sin(th)], 1)              # thin ellipse, closed ring
    rear = np.stack([2.0 + 0.25 * np.cos(th), -0.3 + 0.04 * np.sin(th)], 1)
    outside = lambda ring, c, a, b: ((xz[:, 0] - c[0]) / a) ** 2 + ((xz[:, 1] - c[1]) / b) ** 2 > 1.02
    xz = xz[outside(front, (0.5, 0.0), 0.5, 0.05) & outside(rear, (2.0, -0.3), 0.25, 0.04)]
    field = synth.wake_field(xz, 0, n_t, n_t, components=1, xc=1.0, yc=0.0)[:, 0]  # [N, T]
    metric = field.std(dim=1).to(pt.float64)
    bounds = [[pt.min(xz[:, 0]).item(), pt.min(xz[:, 1]).item()], [pt.max(xz[:, 0]).item(), pt.max(xz[:, 1]).item()]]
    geometry = [CubeGeometry("domain", True, bounds[0], bounds[1]),
                GeometryCoordinates2D("OAT15", False, front, refine=True),
                GeometryCoordinates2D("NACA", False, rear, refine=True)]
    times = pt.arange(n_t, dtype=pt.float64) * 0.5
    s_cube = SparseSpatialSampling(xz, metric, geometry, str(tmp_path), "OAT15_small_area_variance_0.50", "OAT15",
                                   uniform_levels=4, min_metric=0.5, n_jobs=8, max_delta_level=False)
    s_cube.execute_grid_generation()
    nc = s_cube.centers.size(0)
    tree = orc.OracleTree(xz.numpy(), metric.numpy(), geometry, uniform_level=4, min_metric=0.5, sdm_order=1).refine()
    assert np.array_equal(tree.all_centers, s_cube.centers.numpy())
    assert np.array_equal(tree.all_levels, s_cube.levels.numpy())
    # no cell of the grid lies completely inside a body
    for ring, c, a, b in ((front, (0.5, 0.0), 0.5, 0.05), (rear, (2.0, -0.3), 0.25, 0.04)):
        cc = s_cube.centers.numpy()
        half = 0.5 * s_cube.size_initial_cell / 2.0 ** s_cube.levels.numpy().reshape(-1)
        inside_all = np.ones(nc, dtype=bool)
        for sx, sy in ((-1, -1), (-1, 1), (1, 1), (1, -1)):
            px, py = cc[:, 0] + sx * half, cc[:, 1] + sy * half
            inside_all &= ((px - c[0]) / a) ** 2 + ((py - c[1]) / b) ** 2 < 0.95
        assert not inside_all.any()

    export = ExportData(s_cube, write_times=times.tolist())
    export.export(xz, field.unsqueeze(1), "p")
    loader = Dataloader(str(tmp_path), "OAT15_small_area_variance_0.50.h5")
    assert [float(t) for t in loader.write_times] == times.tolist()
    p_s = loader.load_snapshot("p")
    d, idx = orc.knn_search(xz.numpy(), s_cube.centers.numpy(), 8)
    ref_p = orc.interpolate(orc.export_weights(d), idx, field.unsqueeze(1).numpy())[:, 0]
    assert tuple(p_s.shape) == (nc, n_t)
    assert np.abs(p_s.numpy() - ref_p).max() <= 1e-5 * np.abs(field.numpy()).max()
    write_svd_s_cube_to_file("p", str(tmp_path), "OAT15_small_area_variance_0.50", export.new_file, 50, rank=int(1e5))
    out = Dataloader(str(tmp_path), "OAT15_small_area_variance_0.50_p_svd.h5")._store()
    s = np.asarray(out.read("constant/s")).reshape(-1)
    s_ref, _, _ = orc.compute_svd(ref_p.astype(np.float32), loader.weights.numpy(), n_t)
    assert s.shape[0] == n_t and np.abs(s[:4] - s_ref[:4]).max() <= 1e-4 * s_ref[0]
    assert len([k for k in out.keys("constant") if k.startswith("mode_")]) == n_t       # 50 asked, 12 available


def test_export_options(cuda, tmp_path):
    # ExportData options of the reference (export.py:41-126): one file per field, a neighbour count of the caller's
    # choice, and appending a field to an existing file with a second ExportData object
    import synth
    from sparsespatialsampling_b200 import SparseSpatialSampling, ExportData, Dataloader
    from sparsespatialsampling_b200.geometry import CubeGeometry, SphereGeometry
    n_t = 5
    coord = synth.cylinder2d_cloud(3000, seed=61)
    p = synth.wake_field(coord, 0, n_t, n_t, components=1)
    u = synth.wake_field(coord, 0, n_t, n_t, components=2)
    times = [str(i) for i in range(n_t)]
    geoms = [CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
             SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"])]
    s_cube = SparseSpatialSampling(coord, synth.wake_metric(coord), geoms, str(tmp_path), "opt", "grid",
                                   uniform_levels=4, min_metric=0.5)
    s_cube.execute_grid_generation()
    nc = s_cube.centers.size(0)

    def reference(field, k):
        d, idx = orc.knn_search(coord.numpy(), s_cube.centers.numpy(), k)
        return orc.interpolate(orc.export_weights(d), idx, field.numpy())

    # one file per field, 5 neighbours
    export = ExportData(s_cube, write_new_file_for_each_field=True, n_neighbors=5, write_times=times)
    assert export.new_file is True
    export.export(coord, p, "p")
    export.export(coord, u, "U")
    for name, field in (("p", p), ("U", u)):
        assert os.path.exists(tmp_path / f"opt_{name}.xdmf")
        got = Dataloader(str(tmp_path), f"opt_{name}.h5").load_snapshot(name).numpy().reshape(nc, -1, n_t)
        assert np.abs(got - reference(field, 5)).max() <= 1e-5 * float(field.abs().max())
    # default file, then a second object appends another field to it
    export = ExportData(s_cube, write_times=times)
    export.export(coord, p, "p")
    again = ExportData(s_cube, write_times=times, append_existing=True)
    again.export(coord, u, "U")
    loader = Dataloader(str(tmp_path), "opt.h5")
    assert sorted(loader.field_names[times[0]]) == ["U", "p"]
    assert np.abs(loader.load_snapshot("U").numpy() - reference(u, 8)).max() <= 1e-5 * float(u.abs().max())
    assert np.abs(loader.load_snapshot("p").numpy() - reference(p, 8)[:, 0]).max() <= 1e-5 * float(p.abs().max())


def test_svd_from_per_field_files_and_facade_options(cuda, tmp_path):
    # write_svd_s_cube_to_file(new_file=True) reads <name>_<field>.h5 (utils.py:381-386); the facade hands the stopping
    # and scheduling options through to the tree (sparse_spatial_sampling.py:86-115)
    import synth
    from sparsespatialsampling_b200 import (SparseSpatialSampling, ExportData, Dataloader, write_svd_s_cube_to_file)
    from sparsespatialsampling_b200.geometry import CubeGeometry, SphereGeometry
    n_t = 8
    coord = synth.cylinder2d_cloud(3000, seed=71)
    p = synth.wake_field(coord, 0, n_t, n_t, components=1)
    times = [str(i) for i in range(n_t)]
    geoms = [CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
             SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"])]
    metric = synth.wake_metric(coord)
    kw = dict(uniform_levels=3, n_cells_max=400, max_delta_level=True, n_cells_iter_start=25, n_cells_iter_end=5)
    s_cube = SparseSpatialSampling(coord, metric, geoms, str(tmp_path), "f", "grid", **kw)
    s_cube.execute_grid_generation()
    from oracle.topology_oracle import OracleTopology
    ref = orc.OracleTree(coord.numpy(), metric.numpy(), geoms, uniform_level=3, n_cells=400, max_delta_level=True,
                         n_cells_iter_start=25, n_cells_iter_end=5, sdm_order=1, topology=OracleTopology).refine()
    assert np.array_equal(ref.all_centers, s_cube.centers.numpy())
    assert np.array_equal(ref.face_ids, s_cube.faces.numpy()) and np.array_equal(ref.all_nodes, s_cube.vertices.numpy())

    export = ExportData(s_cube, write_new_file_for_each_field=True, write_times=times)
    export.export(coord, p, "p")
    write_svd_s_cube_to_file(["p"], str(tmp_path), "f", export.new_file, 2, rank=3)
    out = Dataloader(str(tmp_path), "f_p_svd.h5")._store()
    assert np.asarray(out.read("constant/s")).reshape(-1).shape[0] == 3
    assert sorted(k for k in out.keys("constant") if k.startswith("mode_")) == ["mode_1", "mode_2"]


def test_export_input_conventions_and_errors(cuda, tmp_path):
    # export.py:154-200: no write times -> ValueError; a 1-D field -> ValueError; a [N, T] field is taken as a scalar
    # field; CUDA tensors (rejected by the reference) are accepted here and give the same result as host tensors
    import synth
    from sparsespatialsampling_b200 import SparseSpatialSampling, ExportData
    from sparsespatialsampling_b200.geometry import CubeGeometry
    coord = synth.cylinder2d_cloud(2000, seed=81)
    p = synth.wake_field(coord, 0, 4, 4, components=1)
    s_cube = SparseSpatialSampling(coord, synth.wake_metric(coord),
                                   [CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"])],
                                   str(tmp_path), "e", "grid", uniform_levels=3, min_metric=0.4)
    s_cube.execute_grid_generation()
    export = ExportData(s_cube, write_files=False)
    with pytest.raises(ValueError):
        export.export(coord, p, "p")                                       # write_times never set
    export.write_times = [str(i) for i in range(4)]
    with pytest.raises(ValueError):
        export.export(coord, p[:, 0, 0], "p")                              # 1-D field
    export.export(coord, p[:, 0, :], "p")                                  # [N, T] -> [N, 1, T]
    host = export.interpolated_fields.centers
    assert tuple(host.shape) == (s_cube.centers.size(0), 1, 4)
    export.export(coord.cuda(), p.cuda(), "p")
    dev = export.interpolated_fields.centers
    assert dev.is_cuda and pt.equal(dev.cpu(), host.cpu())


def test_export_from_a_reloaded_s_cube_object(cuda, tmp_path):
    # examples/s3_for_cylinder3D_Re3900.py:117-122: the pickled s_cube_<name>.pt is loaded again, pointed at a new
    # directory and exported from
    import synth
    from sparsespatialsampling_b200 import SparseSpatialSampling, ExportData, Dataloader
    from sparsespatialsampling_b200.geometry import CubeGeometry
    coord = synth.cylinder2d_cloud(2000, seed=91)
    p = synth.wake_field(coord, 0, 3, 3, components=1)
    first = tmp_path / "first"
    s_cube = SparseSpatialSampling(coord, synth.wake_metric(coord),
                                   [CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"])],
                                   str(first), "r", "grid", uniform_levels=3, min_metric=0.4)
    s_cube.execute_grid_generation()
    direct = ExportData(s_cube, write_times=["0", "1", "2"], write_files=False)
    direct.export(coord, p, "p")
    expected = direct.interpolated_fields.centers.cpu().clone()

    loaded = pt.load(os.path.join(first, "s_cube_r.pt"), weights_only=False)
    second = tmp_path / "second"
    loaded.save_path = str(second)
    export = ExportData(loaded, write_times=["0", "1", "2"])
    export.export(coord, p, "p")
    got = Dataloader(str(second), "r.h5").load_snapshot("p")
    assert pt.equal(got, expected[:, 0])


def test_batchwise_export_loop(cuda, tmp_path):
    # the loop of export_openfoam_fields (utils.py:204-226) with a synthetic reader: ragged last batch, a field that
    # does not exist (reader returns None) is skipped, the file equals a one-shot export
    import synth
    from sparsespatialsampling_b200 import (SparseSpatialSampling, ExportData, Dataloader, export_fields_batchwise,
                                            export_openfoam_fields)
    from sparsespatialsampling_b200.geometry import CubeGeometry
    n_t = 7
    coord = synth.cylinder2d_cloud(2000, seed=101)
    fields = {"p": synth.wake_field(coord, 0, n_t, n_t, components=1), "U": synth.wake_field(coord, 0, n_t, n_t, components=2)}
    times = [str(i) for i in range(n_t)]
    s_cube = SparseSpatialSampling(coord, synth.wake_metric(coord),
                                   [CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"])],
                                   str(tmp_path), "b", "grid", uniform_levels=3, min_metric=0.4)
    s_cube.execute_grid_generation()
    calls = []

    def reader(name, batch_times):
        calls.append((name, list(batch_times)))
        if name not in fields:
            return None, None
        cols = [times.index(t) for t in batch_times]
        return coord, fields[name][:, :, cols]

    export = ExportData(s_cube, write_times=times)
    export_fields_batchwise(export, reader, ["p", "missing", "U"], batch_size=3)
    assert [c[1] for c in calls if c[0] == "p"] == [["0", "1", "2"], ["3", "4", "5"], ["6"]]
    loader = Dataloader(str(tmp_path), "b.h5")
    assert sorted(loader.field_names[times[0]]) == ["U", "p"]
    once = ExportData(s_cube, write_times=times, write_files=False)
    for name in ("p", "U"):
        once.export(coord, fields[name], name)
        expected = once.interpolated_fields.centers.cpu()
        got = loader.load_snapshot(name).reshape(expected.shape)
        assert pt.equal(got, expected), name
    with pytest.raises(ImportError):                                       # flowtorch is not part of this image
        export_openfoam_fields(export, str(tmp_path), [[0, 0], [1, 1]])
