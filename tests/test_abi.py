"""CPU: the C-ABI library loads, exports every symbol include/srba_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "srba_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from srba_stereo_slam_b200 import frontend
    lib = frontend.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_pod_layouts_match_opencv():
    from srba_stereo_slam_b200 import frontend
    assert frontend.KEYPOINT_DTYPE.itemsize == 28 and frontend.DMATCH_DTYPE.itemsize == 16     # cv::KeyPoint / cv::DMatch
    assert list(frontend.KEYPOINT_DTYPE.names) == ["x", "y", "size", "angle", "response", "octave", "class_id"]
    assert list(frontend.DMATCH_DTYPE.names) == ["queryIdx", "trainIdx", "imgIdx", "distance"]
    assert ctypes.sizeof(frontend.SfeConfig) == 17 * 4


def test_default_config_is_kitti_ini():
    from srba_stereo_slam_b200 import frontend
    c = frontend.default_config()
    assert (c.width, c.height, c.nlevels, c.nfeatures) == (1241, 376, 8, 1000)
    assert (c.fast_threshold, c.orb_max_distance, c.edge_threshold) == (20, 60, 31)     # kitti.ini:114,24; ORB default
    assert c.max_y_diff == 2.0 and c.robust_1to1 == 0 and abs(c.scale_factor - 1.2) < 1e-6  # kitti.ini:17,18


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (never route through the oracle or any CPU code)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from srba_stereo_slam_b200 import frontend
    with pytest.raises(frontend.SfeError) as e:
        frontend.StereoFrontEnd()
    assert e.value.code == frontend.SFE_ERR_CUDA
    src = open(os.path.join(ROOT, "srba-stereo-slam_b200", "frontend.py")).read()
    assert "oracle" not in src.replace("never falls back", "")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "srba-stereo-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "frontend_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert "../oracle" not in text and "oracle/" not in text, f


def test_library_is_sm100a_native_with_tma():
    """The shipped cubin is sm_100a only (no fallback architectures) and stages its tiles with TMA
    (UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier), uses DP4A/DP2A and packed 16-bit min/max."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "srba-stereo-slam_b200", "libsrba_b200.so")
    archs = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in archs and not re.search(r"sm_(?!100a)\d+", archs), archs
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert sass.count("UTMALDG") >= 4                       # pyramid (2 modes), FAST, blur (2 modes)
    assert "SYNCS.ARRIVE.TRANS64" in sass and "IDP.4A" in sass and "IDP.2A" in sass
    assert "VIMNMX" in sass and "VABSDIFF4" in sass and "POPC" in sass
    assert "HMMA" not in sass and "UTCHMMA" not in sass     # integer / popc work: tensor cores deliberately unused
