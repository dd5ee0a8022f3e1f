"""ctypes binding of dropin/_build/libslam_dropin[_emu].so: the reference's OWN Frame / KeyFrame / MapPoint / Map / DBoW2 sources
(compiled unmodified from /root/reference) linked with the three DROP-IN translation units of dropin/ -- class ORBextractor,
class OrbMatcher, Frame::ComputeStereoMatches -- behind the same C entry points as oracle/_ref/libslam_ref.so
(oracle/ref_wrap*.cpp).  tests/reference_lib.py is loaded a second time against that library, so a test can run one scenario
through the reference's bodies (reference_lib) and through the GPU bodies (this module) and compare.

kind "gpu": linked against slam_framework_b200/liborbfe.so (needs a B200).  kind "emu": linked against the TEST-ONLY emulated
build of the same kernel sources (tests/emu), for the GPU-less container."""
import importlib.util
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
_mods = {}


def lib_path(kind):
    return os.path.join(ROOT, "dropin", "_build", "libslam_dropin_emu.so" if kind == "emu" else "libslam_dropin.so")


def available(kind):
    return os.path.isdir(os.path.join(REF, "src", "orb_features")) or os.path.exists(lib_path(kind))


def load(kind):
    """a module with the API of tests/reference_lib.py bound to the drop-in library"""
    if kind in _mods:
        return _mods[kind]
    if os.path.isdir(os.path.join(REF, "src", "orb_features")):  # (re)build from the mounted reference tree
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
        if kind == "emu":
            from emu import build_emu
            build_emu.build()
        else:
            from slam_framework_b200 import build as B
            B.build()
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dropin"), "-s"] + (["EMU=1"] if kind == "emu" else []))
    spec = importlib.util.spec_from_file_location("dropin_api_" + kind, os.path.join(HERE, "reference_lib.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.LIB = lib_path(kind)
    m.REF = "/nonexistent"   # lib() must not rebuild oracle/_ref for this binding
    _mods[kind] = m
    return m
