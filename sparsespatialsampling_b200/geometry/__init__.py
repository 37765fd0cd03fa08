from .base import GeometryObject
from .analytic import (CubeGeometry, SphereGeometry, CylinderGeometry3D, TriangleGeometry, PrismGeometry3D,
                       TetrahedronGeometry3D, PyramidGeometry3D)
from .surfaces import GeometrySTL3D, GeometryCoordinates2D

__all__ = ["GeometryObject", "CubeGeometry", "SphereGeometry", "CylinderGeometry3D", "TriangleGeometry",
           "PrismGeometry3D", "TetrahedronGeometry3D", "PyramidGeometry3D", "GeometrySTL3D", "GeometryCoordinates2D"]
