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
