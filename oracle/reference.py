"""
The UNMODIFIED reference (JanisGeise/sparseSpatialSampling v1.0) as a checker and CPU baseline. Test infrastructure:
only tests/, __graft_entry__.smoke() and bench.py's reference arm / cpu_baseline leg may import this module.

`install()` pip-installs the reference from /root/reference (present in the build container only) into the git-ignored
`oracle/_ref/`; the installed package travels to the GPU box with the snapshot (it is not gpurun-ignored), the source
tree under /root/reference does not. No reference source is copied into the repository history.
`load()` returns the imported `sparseSpatialSampling` package (stand-ins of oracle/stubs/ for its absent dependencies)
or None when oracle/_ref/ does not exist.
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def install(force: bool = False) -> bool:
    """pip install --no-index --no-deps --target oracle/_ref <copy of /root/reference>; False if there is no source."""
    marker = os.path.join(REF_DIR, "sparseSpatialSampling", "export.py")
    if os.path.exists(marker) and not force:
        return True
    if not os.path.isdir(REF_SRC):
        return False
    tmp = tempfile.mkdtemp(prefix="s3_ref_src_")
    try:
        src = os.path.join(tmp, "reference")          # the build writes egg-info into the tree: use a copy
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "*.jpeg", "*.png", "docs", "post_processing"))
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", REF_DIR, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return os.path.exists(marker)


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "sparseSpatialSampling", "export.py"))


def load():
    """The reference package, imported from oracle/_ref with the stand-in dependencies; None if not installed."""
    if not available():
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if STUBS not in sys.path:
        sys.path.append(STUBS)                           # last: a real installation of a dependency wins
    pkg = importlib.import_module("sparseSpatialSampling")
    importlib.import_module("sparseSpatialSampling.s_cube")
    importlib.import_module("sparseSpatialSampling.export")
    return pkg


if __name__ == "__main__":
    print("installed" if install(force=True) else "no reference source", REF_DIR)
