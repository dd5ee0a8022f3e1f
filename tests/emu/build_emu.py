"""TEST-ONLY: compiles the CUDA sources of liborbfe (slam_framework_b200/csrc/*.cu) with g++ against
tests/emu/cuda_emu.h so that kernel LOGIC can be diffed against the oracle in the GPU-less build
container.  The result (tests/emu/liborbfe_emu_TESTONLY.so) is never shipped, never loaded by the
package, and reports "EMULATED TEST BUILD" from orbfe_version().  It is not a CPU fallback."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "slam_framework_b200", "csrc")
LIB = os.path.join(HERE, "liborbfe_emu_TESTONLY.so")
SOURCES = ["orbfe_api.cu", "orbfe_match.cu", "orbfe_bow.cu", "orbfe_frame.cu"]


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                                os.path.join(ROOT, "include", "orbfe.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs = []
    for s in SOURCES:
        o = os.path.join(HERE, s + ".emu.o")
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-DORBFE_EMU",
               "-include", os.path.join(HERE, "cuda_emu.h"), "-x", "c++", "-c", os.path.join(CSRC, s), "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
