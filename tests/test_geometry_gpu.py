"""
GPU parity of the geometry masks (GeometryObject.check_cell): the reference's own unit-test triples
(sparseSpatialSampling/tests/test_*_geometry.py, restated) and golden check_cell results of the reference classes on
random and lattice-aligned cells (tests/golden/geometry_pins.npz).
"""
import os

import numpy as np
import pytest
import torch as pt

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class DummyCells:
    """The reference's fixture cells (sparseSpatialSampling/tests/const.py:41-54), float32 on purpose."""
    def __init__(self):
        f = pt.float32
        self.cell_inside_2D = pt.tensor([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=f)
        self.cell_outside_2D = pt.tensor([[5, 5], [6, 5], [6, 6], [5, 6]], dtype=f)
        self.cell_partially_2D = pt.tensor([[0.5, 0.5], [0.5, 1.5], [1.5, 1.5], [1.5, 0.5]], dtype=f)
        self.cell_inside_3D = pt.tensor([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1],
                                         [0, 1, 1]], dtype=f)
        self.cell_outside_3D = pt.tensor([[5, 5, 5], [6, 5, 5], [6, 6, 5], [5, 6, 5], [5, 5, 6], [6, 5, 6], [6, 6, 6],
                                          [5, 6, 6]], dtype=f)
        self.cell_partially_3D = pt.tensor([[0.5, 0.5, 0.5], [1.5, 0.5, 0.5], [1.5, 1.5, 0.5], [0.5, 1.5, 0.5],
                                            [0.5, 0.5, 1.5], [1.5, 0.5, 1.5], [1.5, 1.5, 1.5], [0.5, 1.5, 1.5]], dtype=f)


def _shapes():
    import sparsespatialsampling_b200.geometry as g
    return {
        "cube2d": lambda ki: g.CubeGeometry("c", ki, [0.0, 0.0], [1.0, 1.0]),
        "cube3d": lambda ki: g.CubeGeometry("c", ki, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]),
        "sphere2d": lambda ki: g.SphereGeometry("s", ki, [0.5, 0.5], 0.45),
        "sphere3d": lambda ki: g.SphereGeometry("s", ki, [0.5, 0.5, 0.5], 0.45),
        "cylinder": lambda ki: g.CylinderGeometry3D("cy", ki, [[0.1, 0.2, 0.3], [0.9, 0.7, 0.6]], 0.3),
        "cone": lambda ki: g.CylinderGeometry3D("co", ki, [[0.5, 0.5, 0.0], [0.5, 0.5, 1.0]], [0.4, 0.1]),
        "triangle": lambda ki: g.TriangleGeometry("t", ki, [[0.0, 0.0], [1.0, 0.1], [0.4, 0.9]]),
        "prism": lambda ki: g.PrismGeometry3D("p", ki, [[[0.0, 0.0, 0.1], [1.0, 0.1, 0.1], [0.4, 0.9, 0.1]],
                                                        [[0.0, 0.0, 0.8], [1.0, 0.1, 0.8], [0.4, 0.9, 0.8]]]),
        "tetra": lambda ki: g.TetrahedronGeometry3D("te", ki, [[0.0, 0.0, 0.0], [1.0, 0.1, 0.0], [0.3, 0.9, 0.1],
                                                              [0.4, 0.3, 0.95]]),
        "pyramid": lambda ki: g.PyramidGeometry3D("py", ki, [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0],
                                                            [0.0, 1.0, 0.0], [0.5, 0.5, 1.0]]),
    }


@pytest.mark.parametrize("name", ["cube2d", "cube3d", "sphere2d", "sphere3d", "cylinder", "cone", "triangle", "prism",
                                  "tetra", "pyramid"])
def test_check_cell_matches_reference_golden(cuda, name):
    from sparsespatialsampling_b200.geometry.device import nodes_invalid
    from oracle import s3_oracle as orc
    pins = np.load(os.path.join(GOLDEN, "geometry_pins.npz"))
    nodes, expected = pins[f"{name}_nodes"], pins[f"{name}_invalid"]
    for ci, (ki, rf) in enumerate([(True, False), (False, False), (True, True), (False, True)]):
        g = _shapes()[name](ki)
        got = nodes_invalid([g], pt.from_numpy(nodes), rf).numpy()
        assert np.array_equal(got, expected[:, ci]), (name, ki, rf)
        # and the oracle agrees on the product's own parameter attributes
        assert all(orc.check_cell(g, nodes[t], rf) == expected[t, ci] for t in range(0, nodes.shape[0], 7))


# the reference's unit tests: (shape, keep_inside) x (outside, inside, partially) -> expected check_cell
@pytest.mark.parametrize("shape,dim", [("cube2d", 2), ("cube3d", 3)])
def test_reference_unit_triples_box(cuda, shape, dim):
    # sparseSpatialSampling/tests/test_cube_geometry.py:47-79 (inside cell lies exactly ON the box -> inclusive)
    cells = DummyCells()
    sfx = "2D" if dim == 2 else "3D"
    body, domain = _shapes()[shape](False), _shapes()[shape](True)
    assert body.check_cell(getattr(cells, f"cell_outside_{sfx}")) is False
    assert body.check_cell(getattr(cells, f"cell_inside_{sfx}")) is True
    assert body.check_cell(getattr(cells, f"cell_partially_{sfx}")) is False
    assert domain.check_cell(getattr(cells, f"cell_outside_{sfx}")) is True
    assert domain.check_cell(getattr(cells, f"cell_inside_{sfx}")) is False
    assert domain.check_cell(getattr(cells, f"cell_partially_{sfx}")) is False


def test_reference_unit_triples_other_shapes(cuda):
    import sparsespatialsampling_b200.geometry as g
    cells = DummyCells()
    # test_sphere_geometry.py: r = 0.9 around the cube centre
    for ki, exp in [(False, (False, False, False)), (True, (True, True, False))]:
        s2 = g.SphereGeometry("s", ki, [0.5, 0.5], 0.9)
        got = (s2.check_cell(cells.cell_outside_2D), s2.check_cell(cells.cell_inside_2D),
               s2.check_cell(cells.cell_partially_2D))
        # unit square corners are at distance 0.707 < 0.9 -> all inside: body: invalid; domain: valid
        assert got[0] is (True if ki else False)
        assert got[1] is (False if ki else True)
        assert got[2] is False
    # apply_mask truth table (test_geometry_base.py:124-145)
    cube = g.CubeGeometry("c", False, [0.0, 0.0], [1.0, 1.0])
    assert cube._apply_mask(pt.tensor([True, True, True, True]), False) is True
    assert cube._apply_mask(pt.tensor([True, False, True, True]), False) is False
    dom = g.CubeGeometry("c", True, [0.0, 0.0], [1.0, 1.0])
    assert dom._apply_mask(pt.tensor([False, False, False, False]), False) is True
    assert dom._apply_mask(pt.tensor([False, True, False, False]), False) is False
    assert dom._apply_mask(pt.tensor([False, True, True, True]), True) is True
    assert cube._apply_mask(pt.tensor([False, True, False, False]), True) is True


def test_constructor_assertions(cuda):
    # argument checks of the reference constructors (tests/test_geometry_base.py:95-122 and per-shape tests)
    import sparsespatialsampling_b200.geometry as g
    with pytest.raises(AssertionError):
        g.CubeGeometry("", True, [0.0], [1.0])
    with pytest.raises(AssertionError):
        g.CubeGeometry("c", "yes", [0.0], [1.0])
    with pytest.raises(AssertionError):
        g.CubeGeometry("c", True, [0.0, 0.0], [1.0])
    with pytest.raises(AssertionError):
        g.CubeGeometry("c", True, [1.0, 0.0], [0.0, 1.0])
    with pytest.raises(AssertionError):
        g.SphereGeometry("s", True, [0.0, 0.0], -1.0)
    with pytest.raises(AssertionError):
        g.CylinderGeometry3D("c", True, [[0, 0, 0], [0, 0, 0]], 1.0)
    with pytest.raises(AssertionError):
        g.TriangleGeometry("t", True, [[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]])
    with pytest.raises(AssertionError):
        g.CubeGeometry("c", True, [0.0], [1.0], refine=True, min_refinement_level=0)
    auto = g.CubeGeometry("c", True, [0.0], [1.0], refine=False, min_refinement_level=3)
    assert auto.refine is True
    cyl = g.CylinderGeometry3D("c", False, [[0.1, 0.2, 0.3], [0.9, 0.7, 0.6]], 0.3)
    assert cyl.type == "cylinder" and cyl.center.shape == (3,)
    assert g.SphereGeometry("s", True, [0.0, 0.0], 2).main_width == 2.0
