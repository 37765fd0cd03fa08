"""
Run the REFERENCE's own unit tests against this package (drop-in check of the class surface).

The reference's tests use relative imports (``from ..geometry import CubeGeometry``, ``from ..s_cube import
SamplingTree``), so a throw-away package ``sparseSpatialSampling`` is assembled under /tmp whose sub-modules re-export
this package, with the reference's ``tests`` directory linked in. Nothing is copied into the repository and the test
suite of this repository does not depend on it (the reference tree only exists in the build container).

    python scripts/run_reference_tests.py [/root/reference] [pytest args...]

Geometry tests run on the host; ``test_assignment_*`` drive ``SamplingTree`` and need the GPU.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "/root/reference"
extra = [a for a in sys.argv[1:] if a != ref]
src_tests = os.path.join(ref, "sparseSpatialSampling", "tests")
if not os.path.isdir(src_tests):
    sys.exit(f"reference tests not found under {src_tests}")
shim = tempfile.mkdtemp(prefix="s3_ref_tests_")
pkg = os.path.join(shim, "sparseSpatialSampling")
os.makedirs(os.path.join(pkg, "geometry"))
with open(os.path.join(pkg, "__init__.py"), "w") as f:
    f.write("from sparsespatialsampling_b200 import *\n")
with open(os.path.join(pkg, "geometry", "__init__.py"), "w") as f:
    f.write("from sparsespatialsampling_b200.geometry import *\n")
with open(os.path.join(pkg, "geometry", "geometry_base.py"), "w") as f:
    f.write("from sparsespatialsampling_b200.geometry.base import GeometryObject\n")
for name in ("s_cube", "data", "export", "utils", "sparse_spatial_sampling"):
    with open(os.path.join(pkg, f"{name}.py"), "w") as f:
        f.write(f"from sparsespatialsampling_b200.{name} import *\n"
                f"from sparsespatialsampling_b200.{name} import __dict__ as _d\n" if False else
                f"import sparsespatialsampling_b200.{name} as _m\nglobals().update({{k: v for k, v in vars(_m).items() if not k.startswith('__')}})\n")
shutil.copytree(src_tests, os.path.join(pkg, "tests"))          # throw-away copy under /tmp (fixtures are read relative to it)
env = dict(os.environ, PYTHONPATH=os.pathsep.join([shim, ROOT, os.environ.get("PYTHONPATH", "")]))
rc = subprocess.call([sys.executable, "-m", "pytest", os.path.join(pkg, "tests"), "-q", "-p", "no:cacheprovider"] + extra,
                     env=env, cwd=shim)
shutil.rmtree(shim, ignore_errors=True)
sys.exit(rc)
