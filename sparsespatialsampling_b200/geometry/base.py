"""
Geometry plugin contract of S^3 (host side).

Mirrors ``GeometryObject`` of the reference (sparseSpatialSampling/geometry/geometry_base.py:16-226): same
constructor arguments, same properties (``keep_inside, name, refine, min_refinement_level, type, main_width,
center``) and the same ``check_cell(cell_nodes, refine_geometry=False) -> bool`` meaning (``True`` = the cell is
invalid / touches the surface). The difference is where the test runs: every built-in shape lowers to a parameter
block (``device_params``) that the CUDA mask kernel evaluates for all new cells of a refinement level at once
(``s3_cells_mask``); ``check_cell`` itself evaluates the same device code for a single cell (``s3_nodes_mask``).
"""
import logging
from abc import ABC, abstractmethod

import torch as pt

logger = logging.getLogger(__name__)

# type ids of sparsespatialsampling_b200/csrc/geometry.cuh
GEOM_CUBE, GEOM_SPHERE, GEOM_CYLINDER, GEOM_TRIANGLE, GEOM_PRISM = 0, 1, 2, 3, 4
GEOM_TETRA, GEOM_PYRAMID, GEOM_STL, GEOM_POLY2D, GEOM_CUSTOM = 5, 6, 7, 8, 9


class GeometryObject(ABC):
    def __init__(self, name: str, keep_inside: bool, refine: bool = False, min_refinement_level: int = None):
        self._name = name
        self._keep_inside = keep_inside
        self._refine = refine
        self._min_refinement_level = min_refinement_level
        self._check_common_arguments()

    # ------------------------------------------------------------------ reference behaviour
    def _check_common_arguments(self) -> None:
        # geometry_base.py:78-107
        assert self._name != "", "Found empty string for the geometry object name. Please provide a name."
        assert isinstance(self._keep_inside, bool), (f"Invalid type for argument keep_inside. Expected bool but "
                                                     f"{type(self._keep_inside)} was given.")
        if not self._refine and self._min_refinement_level is not None:
            logger.warning(f"min_refinement_level={self._min_refinement_level} given for geometry {self._name} while "
                           f"refine=False; switching refine on.")
            self._refine = True
        if self._refine and self._min_refinement_level is not None:
            assert self._min_refinement_level > 0, (f"Expected min_refinement_level > 0 but found "
                                                    f"min_refinement_level={self._min_refinement_level}.")

    def _apply_mask(self, mask: pt.Tensor, refine_geometry: bool) -> bool:
        """
        all/any reduction of a per-node inside mask (geometry_base.py:40-76); kept for custom (host) geometries.
        """
        if not refine_geometry:
            invalid = mask.all(0) if not self._keep_inside else ~mask.any(0)
        else:
            invalid = mask.any(0) if not self._keep_inside else ~mask.all(0)
        return bool(invalid.item())

    @property
    def keep_inside(self):
        return self._keep_inside

    @property
    def name(self):
        return self._name

    @property
    def refine(self):
        return self._refine

    @property
    def min_refinement_level(self):
        return self._min_refinement_level

    # ------------------------------------------------------------------ device lowering
    def device_params(self):
        """
        ``(type_id, [fp64 parameters], n_extra)`` for the mask kernel, or ``None`` for a geometry that only has a
        Python ``check_cell`` (evaluated on the host per cell, as the reference does for every geometry).
        """
        return None

    def check_cell(self, cell_nodes: pt.Tensor, refine_geometry: bool = False) -> bool:
        """
        ``True`` if the cell with the given ``[2^d, d]`` nodes is invalid (normal mode) or touches the surface
        (``refine_geometry=True``); evaluated by the CUDA mask code (no CPU fallback).
        """
        from .device import nodes_invalid
        assert cell_nodes.dim() == 2, "cell_nodes must be [n_nodes, n_dimensions]"
        self._check_dimensions(cell_nodes)
        return bool(nodes_invalid([self], cell_nodes.unsqueeze(0), refine_geometry)[0])

    def _check_dimensions(self, cell_nodes: pt.Tensor) -> None:
        pass

    @abstractmethod
    def _check_geometry(self) -> None:
        pass

    @property
    @abstractmethod
    def type(self) -> str:
        pass

    @property
    @abstractmethod
    def main_width(self) -> float:
        pass

    @property
    @abstractmethod
    def center(self) -> pt.Tensor:
        pass
