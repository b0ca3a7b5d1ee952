"""The C-ABI library loads without a GPU and exports every symbol include/velocity_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "velocity_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vel_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header():
    from velocity_b200 import _lib, build

    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert not _lib.PENDING


def test_host_only_entry_points():
    from velocity_b200 import _lib

    assert _lib.lib().vel_version() == 100
    lay = _lib.pyr_layout(1920, 1080, (15, 15), 4)
    assert lay.max_level == 4 and list(lay.width[:5]) == [1920, 960, 480, 240, 120]
    assert list(lay.height[:5]) == [1080, 540, 270, 135, 68]
    lay = _lib.pyr_layout(320, 240, (15, 15), 4)
    assert lay.max_level == 3                       # OpenCV stops before a level with height <= win
    with pytest.raises(RuntimeError):
        _lib.pyr_layout(0, 10, (15, 15), 2)
    assert b"bad size" in _lib.lib().vel_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "velocity_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "from oracle" not in txt and "import oracle" not in txt, fn


def test_shipped_library_links_no_vendor_blas_or_solver():
    """K8 is hand-written (csrc/dense_f64.cu): the default build must not depend on cuBLAS / cuSOLVER."""
    import subprocess

    from velocity_b200 import _lib, build

    build.build()
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout.lower()
    sym = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout.lower()
    for name in ("cublas", "cusolver", "cusparse", "cufft"):
        assert name not in out and name not in sym, name
