"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU), exports every symbol include/mrslam_b200.h
declares, and refuses to compute without a device instead of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from mr_slam_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mrslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrslam_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) > 50
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, "declared in include/mrslam_b200.h but not exported: %s" % missing
    assert b"sm_100a" in L.mrslam_version()


def test_header_compiles_as_plain_c(tmp_path):
    """the boundary is a C ABI: no C++ or torch types in the signatures"""
    src = tmp_path / "t.c"
    src.write_text('#include "mrslam_b200.h"\nint main(void){ return MRSLAM_OK; }\n')
    import subprocess
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU runner")
    L = _lib.lib()
    assert L.mrslam_init(0) != 0
    assert b"no CPU fallback" in L.mrslam_last_error() or b"CUDA" in L.mrslam_last_error()
    assert not L.mrslam_gicp_create()          # NULL handle, error string set
    assert not L.mrslam_elev_create()
    pts = np.zeros(30, np.float32)
    out = np.zeros(120 * 120 * 3, np.float32)
    assert L.mrslam_voxelocc_transform(pts.ctypes.data, 10, 1, 1, 120, 120, 1, 1, out.ctypes.data) != 0
    from mr_slam_b200 import pygicp
    with pytest.raises(_lib.MrSlamError):
        pygicp.FastGICP()


def test_product_never_touches_the_oracle():
    """the CUDA library and the host mirrors must not import, link or load anything under oracle/"""
    pkg = os.path.join(ROOT, "mr_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
    import subprocess
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
