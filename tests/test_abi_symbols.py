"""The C-ABI library builds with nvcc (sm_100a cross-compile, no GPU needed), loads, and exports
every entry point include/orbfe.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

from slam_framework_b200 import build as B
from slam_framework_b200 import orbfe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "orbfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orbfe_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list():
    assert declared_symbols() == sorted(orbfe.EXPORTS)


def test_library_builds_and_exports_all_symbols():
    lib = ctypes.CDLL(B.build())
    for s in declared_symbols():
        assert hasattr(lib, s), f"liborbfe.so does not export {s}"


def test_sm100a_cubin_embedded():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", B.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    """On a box with no CUDA device every compute entry point must fail loudly."""
    L = orbfe.load()
    if L.orbfe_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(orbfe.OrbfeError):
        orbfe.ORBextractor(lib=L)


def test_product_loader_refuses_emulated_build():
    from emu import build_emu
    with pytest.raises(orbfe.OrbfeError):
        orbfe.load(build_emu.build())


def test_every_stateless_entry_point_fails_loudly_without_device():
    """frame / vocabulary / frame-tail entry points: no CUDA device => an error, never a CPU result"""
    import numpy as np
    L = orbfe.load()
    if L.orbfe_device_count() > 0:
        pytest.skip("CUDA device present")
    kps = np.zeros(4, orbfe.KP_DTYPE)
    desc = np.zeros((4, 32), np.uint8)
    scale = np.array([1.0, 1.2], np.float32)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.Frame(kps, desc, scale, (0, 100, 0, 100), lib=L)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.DescriptorDistance(desc, desc, lib=L)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.OrbVocabulary(2, 1, 0, 0, [0, 0, 0], [0, 1, 1], np.zeros((3, 32), np.uint8), [0.0, 1.0, 1.0], lib=L)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.UndistortKeyPoints(kps, 500.0, 500.0, 320.0, 240.0, [0.1, 0.0, 0.0, 0.0], lib=L)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.IsInFrustum(np.ones((2, 3), np.float32), np.ones((2, 3), np.float32), [0.1, 0.1], [9.0, 9.0], [8.0, 8.0], np.eye(3),
                          np.zeros(3), np.zeros(3), 500.0, 500.0, 320.0, 240.0, 40.0, (0, 640, 0, 480), 0.18, 8, lib=L)
    with pytest.raises(orbfe.OrbfeError):
        orbfe.debug_logf(np.ones(4, np.float32), lib=L)


def test_argument_validation_precedes_everything():
    """bad arguments are rejected with a message (no device needed to see it)"""
    import numpy as np
    L = orbfe.load()
    with pytest.raises(orbfe.OrbfeError, match="vocabulary header"):
        orbfe.OrbVocabulary(25, 6, 0, 0, [0], [0], np.zeros((1, 32), np.uint8), [0.0], lib=L)          # k > 20 (TemplatedVocabulary.h:1356)
    with pytest.raises(orbfe.OrbfeError, match="precede"):
        orbfe.OrbVocabulary(2, 1, 0, 0, [0, 2, 0], [0, 1, 1], np.zeros((3, 32), np.uint8), [0.0, 1.0, 1.0], lib=L)
    with pytest.raises(orbfe.OrbfeError, match="tilted"):
        orbfe.UndistortKeyPoints(np.zeros(1, orbfe.KP_DTYPE), 500.0, 500.0, 320.0, 240.0, [0.1] * 12 + [0.01, 0.0], lib=L)
    with pytest.raises(orbfe.OrbfeError, match="vocabulary text file|cannot open"):
        orbfe.OrbVocabulary.loadFromTextFile("/nonexistent/ORBvoc.txt", lib=L)
