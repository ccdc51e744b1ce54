"""CPU: the C-ABI library loads and exports every symbol include/lsdreg.h declares; without a
device every entry point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import lsdreg
from lsdreg import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lsdreg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 30
    bound = {n for n, _, _ in capi.SIGNATURES}
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in lsdreg.h but not exported by liblsdreg.so"
        assert n in bound, f"{n} has no ctypes signature in capi.py"
    assert bound <= set(names), f"bound but undeclared: {bound - set(names)}"


def test_header_cites_reference_for_each_group():
    src = open(os.path.join(ROOT, "include", "lsdreg.h")).read()
    for cite in ("ivox3d.h:", "laserMapping.cpp:", "esekfom.hpp:", "ikd_Tree.cpp:", "use-ikfom.hpp:"):
        assert cite in src


def test_product_never_touches_the_oracle():
    """No import / include / dlopen of anything under oracle/ from the product tree."""
    pkg = os.path.join(ROOT, "lidar-slam-detection_b200")
    py_imp = re.compile(r"^\s*(import|from)\s+oracle\b", re.M)
    c_inc = re.compile(r"#\s*include\s*[<\"][^>\"]*oracle|dlopen\s*\([^)]*oracle|liblsd_oracle|libref_(lio|reg|cuda|vfe|ikfom|fastlio|keyframe)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            txt = None
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not py_imp.search(txt), f
                assert "liblsd_oracle" not in txt and "libref_" not in txt, f
            elif f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not c_inc.search(txt), f


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a host without a CUDA device")
def test_fails_loudly_without_a_device():
    with pytest.raises(lsdreg.LsdError) as e:
        lsdreg.HashVoxelMap(0.5, 12)
    assert e.value.status == lsdreg.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(lsdreg.LsdError):
        lsdreg.LioFrontend()
    with pytest.raises(lsdreg.LsdError):
        lsdreg.VoxelGrid(100)
    with pytest.raises(lsdreg.LsdError):
        lsdreg.ImuProcess()
    with pytest.raises(lsdreg.LsdError):
        lsdreg.Matcher("FAST_VGICP")
    with pytest.raises(lsdreg.LsdError) as e:
        lsdreg.ScanContext(db_capacity=16)
    assert e.value.status == lsdreg.ERR_NO_DEVICE


def test_host_only_helpers_work_anywhere():
    P = lsdreg.init_cov()
    assert P.shape == (23, 23) and P[0, 0] == 1.0 and P[6, 6] == 1e-5   # IMU_Processing.hpp:224-230
    assert capi.lib.lsd_version().startswith(b"lsdreg")


def test_slam_wrapper_module_has_the_reference_entry():
    """The pybind11 module keeps the reference's name, function name and argument names
    (slam_wrapper.cpp:241-242) and, like every entry point, refuses to run without a GPU."""
    import importlib
    import inspect
    import os
    import sys
    import numpy as np
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lidar-slam-detection_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    slam = importlib.import_module("slam_wrapper")
    doc = slam.pointcloud_align.__doc__
    for name in ("source_point", "target_point", "guess"):
        assert name in doc
    import torch
    if not torch.cuda.is_available():
        import pytest
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            slam.pointcloud_align(np.zeros((10, 4), np.float32), np.zeros((10, 4), np.float32), np.eye(4, dtype=np.float32))
