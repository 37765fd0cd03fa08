"""CPU tests of the I/O format (Datawriter / Dataloader / XDMFWriter) against outputs of the reference's own classes
(tests/golden/io_golden.npz, written by tests/golden/make_golden_io.py) and of the SVD oracle against the reference's
compute_svd."""
import os

import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "io_golden.npz"))


@pytest.fixture(scope="module")
def grid():
    return np.load(os.path.join(GOLDEN, "g2d_metric.npz"))


def write_case(tmp, gold, grid, with_metric=True):
    from sparsespatialsampling_b200.data import Datawriter
    w = Datawriter(tmp, "case.h5")
    w.write_data("faces", group="grid", data=pt.from_numpy(grid["faces"]))
    w.write_data("vertices", group="grid", data=pt.from_numpy(grid["vertices"]))
    w.write_data("centers", group="grid", data=pt.from_numpy(grid["centers"]))
    w.write_data("levels", group="constant", data=pt.from_numpy(grid["levels"]))
    if with_metric:
        w.write_data("metric", group="constant", data=pt.from_numpy(gold["p"][:, 0].astype(np.float64)))
    w.write_data("size_initial_cell", group="constant", data=float(grid["width"]))
    for i, t in enumerate(gold["times"].tolist()):
        w.write_data("p_center", group="data", time_step=t, data=pt.from_numpy(gold["p"][:, i]))
        w.write_data("U_center", group="data", time_step=t, data=pt.from_numpy(gold["U"][:, :, i]))
    w.close()
    w.write_xdmf_file()


def test_temporal_xdmf_is_byte_identical_to_the_reference(tmp_path, gold, grid):
    write_case(str(tmp_path), gold, grid)
    assert open(tmp_path / "case.xdmf").read() == str(gold["xdmf_temporal"])


def test_dataloader_matches_reference(tmp_path, gold, grid):
    from sparsespatialsampling_b200.data import Dataloader
    write_case(str(tmp_path), gold, grid)
    loader = Dataloader(str(tmp_path), "case.h5")
    assert loader.write_times == gold["write_times"].tolist()
    assert np.array_equal(loader.weights.numpy(), gold["weights"])
    assert np.array_equal(loader.load_snapshot("p").numpy(), gold["dm_p"])
    assert np.array_equal(loader.load_snapshot("U").numpy(), gold["dm_U"])
    both = loader.load_snapshot(["p", "U"], gold["times"].tolist()[:2])
    assert tuple(both[0].shape) == (gold["p"].shape[0], 2) and tuple(both[1].shape) == (gold["p"].shape[0], 2, 2)
    assert sorted(loader.field_names["0.5"]) == ["U", "p"]
    assert np.array_equal(loader.vertices.numpy(), grid["centers"]) and np.array_equal(loader.faces.numpy(), grid["faces"])
    assert np.array_equal(loader.levels.numpy(), grid["levels"].squeeze())


def test_constant_xdmf_is_byte_identical_to_the_reference(tmp_path, gold, grid):
    from sparsespatialsampling_b200.data import Datawriter, Dataloader
    write_case(str(tmp_path), gold, grid)
    loader = Dataloader(str(tmp_path), "case.h5")
    w = Datawriter(str(tmp_path), "case_p_svd.h5")
    w.write_grid(loader)
    for i in range(3):
        w.write_data(f"mode_{i + 1}", group="constant", data=pt.from_numpy(gold["svd_p_U"][:, i]))
    w.write_data("V", group="constant", data=pt.from_numpy(gold["svd_p_V"]))
    w.write_data("s", group="constant", data=pt.from_numpy(gold["svd_p_s"]))
    w.write_data("cell_area", group="constant", data=loader.weights)
    w.write_xdmf_file()
    assert open(tmp_path / "case_p_svd.xdmf").read() == str(gold["xdmf_const"])


def test_append_mode_and_duplicate_fields(tmp_path, gold, grid):
    from sparsespatialsampling_b200.data import Datawriter, Dataloader
    write_case(str(tmp_path), gold, grid)
    w = Datawriter(str(tmp_path), "case.h5", mode="a")
    w.write_data("T_center", group="data", time_step="0.1", data=pt.zeros(gold["p"].shape[0]))
    w.write_data("p_center", group="data", time_step="0.1", data=pt.zeros(gold["p"].shape[0]))   # exists: skipped
    with pytest.raises(ValueError):
        w.write_data("x", group="nonsense", data=pt.zeros(3))
    w.close()
    loader = Dataloader(str(tmp_path), "case.h5")
    assert "T" in loader.field_names["0.1"]
    assert np.array_equal(loader.load_snapshot("p", "0.1").numpy()[:, 0], gold["p"][:, 0])


@pytest.mark.parametrize("name", ["p", "U"])
def test_svd_oracle_matches_reference_compute_svd(gold, name):
    s, u, v = orc.compute_svd(gold[f"dm_{name}"], gold["weights"].astype(np.float32), rank=4)
    assert np.allclose(s, gold[f"svd_{name}_s"], rtol=2e-5)
    assert u.shape == gold[f"svd_{name}_U"].shape and v.shape == gold[f"svd_{name}_V"].shape
    for i in range(3):
        assert abs(np.dot(v[:, i], gold[f"svd_{name}_V"][:, i])) > 1 - 1e-4
        a, b = u[..., i].ravel(), gold[f"svd_{name}_U"][..., i].ravel()
        assert abs(np.dot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b)) > 1 - 1e-4


def test_cell_area_oracle(gold, grid):
    assert np.allclose(orc.cell_area(float(grid["width"]), grid["levels"], 2), gold["weights"], rtol=1e-15)
