"""
CPU tests of the built-in HDF5 layer (sparsespatialsampling_b200/h5lite.py).

Pins: the reference's own golden output file (sparseSpatialSampling/tests/s_cube_test_dataset.h5, written by h5py;
committed as a binary fixture) is read with h5lite and through the product ``Dataloader`` with the assertions of the
reference's tests/test_s_cube_dataloader.py:23-57; files written by h5lite are read back, byte-level structure checks
follow the HDF5 file format specification (superblock v0, symbol-table groups).
"""
import os
import struct

import numpy as np
import pytest
import torch as pt

from sparsespatialsampling_b200 import h5lite
from sparsespatialsampling_b200.data import Dataloader, Datawriter

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_FILE = "reference_s_cube_test_dataset.h5"


def test_reads_the_reference_golden_file():
    f = h5lite.File(os.path.join(GOLD, REF_FILE), "r")
    assert f.keys() == ["constant", "data", "grid"]
    assert f.keys("grid") == ["centers", "faces", "vertices"] and f.keys("data") == ["0.4"]
    assert f.read("grid/centers").shape == (209, 2) and f.read("grid/centers").dtype == np.float64
    assert f.read("grid/faces").shape == (209, 4) and f.read("grid/faces").dtype == np.int32
    assert f.read("grid/vertices").shape == (247, 2)
    assert f.read("constant/levels").dtype == np.int64 and f.read("constant/size_initial_cell").shape == ()
    faces = f.read("grid/faces")
    assert faces.min() >= 0 and faces.max() < 247          # a consistent topology came out of the bytes
    f.close()


def test_product_dataloader_on_the_reference_file():
    # the reference's own test (tests/test_s_cube_dataloader.py:39-57), against the product Dataloader
    n_cells, n_nodes, n_dimensions = 209, 247, 2
    loader = Dataloader(GOLD, REF_FILE)
    assert len(loader.write_times) == 1 and loader.write_times == ["0.4"]
    assert loader.field_names == {"0.4": ["p"]}
    assert loader.vertices.shape == (n_cells, n_dimensions)
    assert loader.weights.shape == loader.levels.shape
    assert loader.faces.shape == (n_cells, pow(2, n_dimensions))
    assert loader.nodes.shape == (n_nodes, n_dimensions)
    assert loader.load_snapshot("p", "0.4").shape == (n_cells, 1)
    # cell areas follow from levels and the root size (data.py:240-247)
    w = loader.weights
    assert float(w.min()) > 0 and abs(float(w.max()) / float(w.min()) - 4.0 ** int(loader.levels.max() - loader.levels.min())) < 1e-9


def test_round_trip_dtypes_shapes_and_groups(tmp_path):
    rng = np.random.default_rng(0)
    arrays = {
        ("grid", "centers"): rng.random((37, 3)), ("grid", "faces"): rng.integers(0, 99, (37, 8)).astype(np.int32),
        ("constant", "levels"): rng.integers(0, 9, (37, 1)), ("constant", "size_initial_cell"): np.float64(2.2),
        ("constant", "flag"): np.arange(5, dtype=np.uint8), ("constant", "empty"): np.zeros((0, 4), dtype=np.float32),
        ("data/0.25", "p_center"): rng.random(37).astype(np.float32),
        ("data/0.25", "U_center"): rng.random((37, 3)).astype(np.float32),
        ("deep/er/still", "x"): rng.integers(-5, 5, (2, 3, 4)).astype(np.int16),
    }
    path = str(tmp_path / "t.h5")
    with h5lite.File(path, "w") as f:
        for (g, n), a in arrays.items():
            assert f.write(g, n, a)
        assert not f.write("grid", "centers", np.zeros(3))          # exists: nothing written
    with h5lite.File(path, "r") as f:
        assert f.keys() == ["constant", "data", "deep", "grid"]
        for (g, n), a in arrays.items():
            got = f.read(f"{g}/{n}")
            assert got.dtype == np.asarray(a).dtype and got.shape == np.shape(a) and np.array_equal(got, a)
        with pytest.raises(KeyError):
            f.read("grid/nope")
    raw = open(path, "rb").read()
    assert raw[:8] == h5lite.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)         # end-of-file address


def test_many_links_need_several_btree_levels_and_append_mode(tmp_path):
    # 1000 time-step groups (a C2 export): symbol nodes of 8 links, B-tree nodes of 32 children -> three levels
    path = str(tmp_path / "big.h5")
    times = [f"{0.001 * i:.3f}" for i in range(1000)]
    with h5lite.File(path, "w") as f:
        f.write("grid", "centers", np.zeros((4, 2)))
        for t in times[:600]:
            f.write(f"data/{t}", "p_center", np.full(4, float(t), dtype=np.float32))
    size_first = os.path.getsize(path)
    with h5lite.File(path, "a") as f:                               # batch 2 appends, nothing is rewritten
        assert len(f.keys("data")) == 600
        for t in times[600:]:
            f.write(f"data/{t}", "p_center", np.full(4, float(t), dtype=np.float32))
            f.write(f"data/{t}", "U_center", np.full((4, 2), float(t), dtype=np.float32))
    assert os.path.getsize(path) > size_first
    with h5lite.File(path, "r") as f:
        assert f.keys("data") == sorted(times, key=lambda s: s.encode())
        for t in (times[0], times[599], times[600], times[999]):
            assert np.array_equal(f.read(f"data/{t}/p_center"), np.full(4, float(t), dtype=np.float32))
        assert f.keys(f"data/{times[999]}") == ["U_center", "p_center"]
    # structure: every B-tree node's keys are ascending names and bracket its children
    rd = h5lite._Reader(open(path, "rb"))
    msgs = rd.messages(rd.root_header)
    bt, hp = struct.unpack_from("<QQ", [d for t, d in msgs if t == h5lite.MSG_SYMBOL_TABLE][0], 0)
    names = [n for n, _ in rd.symbols(bt, rd.heap_names(hp))]
    assert names == ["data", "grid"]


def test_datawriter_files_are_real_hdf5_and_reload(tmp_path):
    w = Datawriter(str(tmp_path), "case.h5")
    centers = pt.rand(10, 2, dtype=pt.float64)
    w.write_data("centers", group="grid", data=centers)
    w.write_data("vertices", group="grid", data=pt.rand(18, 2, dtype=pt.float64))
    w.write_data("faces", group="grid", data=pt.randint(0, 18, (10, 4), dtype=pt.int32))
    w.write_data("levels", group="constant", data=pt.randint(1, 4, (10, 1)))
    w.write_data("size_initial_cell", group="constant", data=1.5)
    w.write_data("p_center", group="data", time_step="0.1", data=pt.rand(10))
    w.write_xdmf_file()
    assert open(os.path.join(tmp_path, "case.h5"), "rb").read(8) == h5lite.SIGNATURE
    assert not os.path.exists(os.path.join(tmp_path, "case.h5.pt"))
    loader = Dataloader(str(tmp_path), "case.h5")
    assert pt.equal(loader.vertices, centers) and loader.write_times == ["0.1"] and loader.field_names == {"0.1": ["p"]}
    assert "case.h5:/data/0.1/p_center" in open(os.path.join(tmp_path, "case.xdmf")).read()
