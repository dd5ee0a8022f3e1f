"""Builds liborbfe.so (the C-ABI CUDA library of include/orbfe.h) in-tree with nvcc for sm_100a.

    python -m slam_framework_b200.build [--force]

The library is the product: hand-written CUDA kernels + a thin C++ host layer.  There is no CPU
build of it; nvcc cross-compiles without a GPU, running it needs a B200.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liborbfe.so")
SOURCES = ["orbfe_api.cu", "orbfe_match.cu", "orbfe_bow.cu", "orbfe_frame.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # IEEE parity with the reference's CPU path: no fused multiply-add
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: liborbfe.so cannot be built (there is no CPU build of this library)")


def _deps():
    out = [os.path.join(ROOT, "include", "orbfe.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h", ".inc")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, out=None, defines=()):
    """Compile liborbfe.so if sources are newer; returns its path.  `out`/`defines` build a tuning
    variant (-DNAME=VALUE) next to the product library for A/B runs (ORBFE_LIB=...)."""
    global LIB
    if out is None and not force and not needs_build():
        return LIB
    target = out or LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          ["-o", target + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building liborbfe.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    os.replace(target + ".tmp", target)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
