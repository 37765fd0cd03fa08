"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol of include/s3b200.h."""
import os
import re

import pytest
import torch as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "s3b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s3_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sparsespatialsampling_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/s3b200.h but not exported by libs3b200.so"
    # the ctypes table mirrors the header one to one
    assert sorted(_lib.exported_symbols()) == declared
    assert lib.s3_version() >= 100
    assert isinstance(lib.s3_last_error(), bytes)


def test_no_cpu_fallback_without_a_device():
    # the product path must fail loudly when there is no CUDA device (no oracle / CPU fallback behind it)
    if pt.cuda.is_available():
        pytest.skip("CUDA device present")
    from sparsespatialsampling_b200 import _lib, KnnIndex, interpolate_data
    import sparsespatialsampling_b200.geometry as g
    with pytest.raises(_lib.S3Error):
        KnnIndex(pt.rand(10, 2, dtype=pt.float64))
    with pytest.raises(_lib.S3Error):
        interpolate_data(pt.rand(4, 2), pt.zeros(4, 2, dtype=pt.int64), pt.rand(5, 1, 3))
    with pytest.raises(_lib.S3Error):
        g.CubeGeometry("c", True, [0.0, 0.0], [1.0, 1.0]).check_cell(pt.zeros(4, 2))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparsespatialsampling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", "").lower() or \
                    "import" not in "".join(l for l in src.splitlines() if "oracle" in l.lower()), f
