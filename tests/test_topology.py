"""
CPU tests of the host-side cell topology (csrc/topology.cu via sparsespatialsampling_b200.topology.Topology) and of its
oracle (oracle/topology_oracle.py):
  * the known answers of the reference's own unit tests (sparseSpatialSampling/tests/test_assignment_neighbors.py,
    test_assignment_nodes.py: uniform grids on [0,10]^d),
  * faces / vertices of reference runs (tests/golden/*.npz), reached by replaying the refinement history with the CPU
    oracle tree -- including the two cases where the reference's neighbour pointers go stale (g3d_delta, g2d_delta_geo).
"""
import os

import numpy as np
import pytest

from oracle import s3_oracle as orc
from oracle.topology_oracle import OracleTopology, neighbour_table, slot_directions
from tests.golden.make_golden import case_definitions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["g2d_metric", "g2d_ncells", "g3d_metric", "g2d_delta", "g3d_delta", "g2d_delta_geo", "g2d_reltol",
         "g3d_delta_geo"]


def _product_topology(d, center, width):
    from sparsespatialsampling_b200.topology import Topology
    return Topology(d, center, width)


def _product_topology_async(d, center, width):
    from sparsespatialsampling_b200.topology import Topology
    return Topology(d, center, width, asynchronous=True)


FACTORIES = {"oracle": OracleTopology, "product": _product_topology}


def _uniform(factory, d, levels):
    t = factory(d, np.full(d, 5.0), 10.0)
    leaves = [0]
    for _ in range(levels):
        first = 1 if leaves == [0] else leaves[-1] + 1
        t.refine(leaves)
        t.refresh_children(leaves)
        leaves = list(range(first, first + len(leaves) * 2 ** d))
    return t


def _cell(t, c):
    if isinstance(t, OracleTopology):
        return list(t.nb[c]), list(t.node[c])
    nb, node, _, _, _ = t.cell(c)
    return nb.tolist(), node.tolist()


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_reference_unit_test_pins_2d(impl):
    # test_assignment_neighbors.py:11-119 and test_assignment_nodes.py:11-85 (two uniform levels, 4 x 4 grid)
    t = _uniform(FACTORIES[impl], 2, 2)
    nb_pins = {5: [-1, -1, 6, 7, 8, -1, -1, -1], 7: [6, 9, 12, 13, 18, 17, 8, 5], 13: [12, 11, 14, 15, 16, 19, 18, 7],
               10: [-1, -1, -1, -1, 11, 12, 9, -1], 15: [14, -1, -1, -1, -1, -1, 16, 13],
               20: [17, 18, 19, -1, -1, -1, -1, -1]}
    node_pins = {5: [0, 9, 10, 11], 7: [10, 12, 5, 13], 12: [12, 15, 17, 5], 13: [5, 17, 18, 19], 15: [18, 20, 2, 21]}
    for c, want in nb_pins.items():
        assert _cell(t, c)[0] == want, c
    for c, want in node_pins.items():
        assert _cell(t, c)[1] == want, c


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_reference_unit_test_pins_3d(impl):
    # test_assignment_nodes.py:88-198 (one level: 27 nodes) and test_assignment_neighbors.py:122-212 (two levels)
    t1 = _uniform(FACTORIES[impl], 3, 1)
    n_nodes = len(t1.nodes) if isinstance(t1, OracleTopology) else t1.n_nodes
    assert n_nodes == 27
    node_pins = {1: [0, 8, 9, 10, 11, 12, 13, 14], 2: [8, 1, 15, 9, 12, 16, 17, 13], 3: [9, 15, 2, 18, 13, 17, 19, 20],
                 4: [10, 9, 18, 3, 14, 13, 20, 21], 5: [11, 12, 13, 14, 4, 22, 23, 24], 6: [12, 16, 17, 13, 22, 5, 25, 23],
                 7: [13, 17, 19, 20, 23, 25, 6, 26], 8: [14, 13, 20, 21, 24, 23, 26, 7]}
    for c, want in node_pins.items():
        assert _cell(t1, c)[1] == want, c
    t2 = _uniform(FACTORIES[impl], 3, 2)
    nb9 = _cell(t2, 9)[0]
    assert nb9[8:17] == [-1, -1, 14, 15, 16, -1, -1, -1, 13] and nb9[17:] == [-1] * 9
    nb43 = _cell(t2, 43)[0]
    assert nb43[8:17] == [46, 53, 56, 61, 70, 69, 48, 45, 47] and nb43[17:] == [14, 21, 24, 29, 38, 37, 16, 13, 15]
    nb32 = _cell(t2, 32)[0]
    assert nb32[8:17] == [57, 58, 59, -1, -1, -1, 67, 66, 60] and nb32[17:] == [25, 26, 27, -1, -1, -1, 35, 34, 28]


@pytest.mark.parametrize("d", [2, 3])
def test_neighbour_table_is_geometrically_consistent(d):
    table, dirs = neighbour_table(d), slot_directions(d)
    assert len(table) == 2 ** d and all(len(row) == len(dirs) for row in table)
    # every sibling appears exactly once per child; entries leaving the parent name the opposite-facing child
    for c, row in enumerate(table):
        assert sorted(e[1] for e in row if e[0] == "sibling") == [j for j in range(2 ** d) if j != c]
        assert sum(e[0] == "poc" for e in row) == len(dirs) - (2 ** d - 1)


@pytest.mark.parametrize("impl", ["oracle", "product"])
@pytest.mark.parametrize("name", CASES)
def test_replay_reproduces_reference_faces_and_vertices(name, impl):
    import sparsespatialsampling_b200.geometry as geo
    case = case_definitions(geo)[name]
    ref = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    tree = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"],
                          sdm_order=1, topology=FACTORIES[impl]).refine()
    assert tree.leaf_order == ref["leaf_index"].tolist()
    assert np.array_equal(tree.all_centers, ref["centers"])
    assert tree.face_ids.dtype == np.int32 and np.array_equal(tree.face_ids, ref["faces"])
    assert np.array_equal(tree.all_nodes, ref["vertices"])
    assert tree.n_cells_log == ref["cells_per_iter"].tolist()


def test_product_and_oracle_keep_identical_pointers():
    import sparsespatialsampling_b200.geometry as geo
    case = case_definitions(geo)["g3d_delta"]
    trees = {k: orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"],
                               sdm_order=1, topology=f) for k, f in FACTORIES.items()}
    for t in trees.values():
        t.refine()
    a, b = trees["oracle"].topo, trees["product"].topo
    assert len(a.parent) == b.n_cells and len(a.nodes) == b.n_nodes
    for c in range(b.n_cells):
        nb, node, parent, children, level = b.cell(c)
        assert nb.tolist() == a.nb[c] and node.tolist() == a.node[c]
        assert (parent, children, level) == (a.parent[c], a.children[c], a.level[c])


def test_stale_pointer_cases_differ_from_the_geometric_closure():
    # documents why the pointer replay exists: with max_delta_level + geometry refinement the leaf covering the
    # adjacent lattice position is not what the reference's pointers refer to
    import sparsespatialsampling_b200.geometry as geo
    case = case_definitions(geo)["g2d_delta_geo"]
    ref = np.load(os.path.join(GOLDEN, "g2d_delta_geo.npz"))
    tree = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"],
                          sdm_order=1).refine()
    assert tree.leaf_order != ref["leaf_index"].tolist() or not np.array_equal(tree.all_centers, ref["centers"])


def test_asynchronous_replay_gives_the_same_grid():
    # updates queued on the library's worker thread (the mode the refinement loop uses without max_delta_level)
    import sparsespatialsampling_b200.geometry as geo
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.topology import Topology
    for name in ("g2d_ncells", "g3d_metric"):
        case = case_definitions(geo)[name]
        ref = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        tree = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"],
                              sdm_order=1, topology=_product_topology_async).refine()
        assert np.array_equal(tree.face_ids, ref["faces"]) and np.array_equal(tree.all_nodes, ref["vertices"])
    t = Topology(2, [0.0, 0.0], 1.0, asynchronous=True)
    t.refine([0])
    t.refine([0])                          # queued: the failure surfaces at the next synchronisation point
    with pytest.raises(_lib.S3Error):
        t.sync()


def test_topology_errors():
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.topology import Topology
    t = Topology(2, [0.0, 0.0], 1.0)
    t.refine([0])
    with pytest.raises(_lib.S3Error):
        t.refine([0])                      # already has children
    with pytest.raises(_lib.S3Error):
        t.mark_invalid([99])
    with pytest.raises(_lib.S3Error):
        t.check_nb(-3)
    faces, vertices, centers = t.final()
    assert faces.shape == (4, 4) and vertices.shape == (9, 2) and centers.shape == (5, 2)
