"""Bounds checking of every kernel's memory accesses on CPU: the emulated TEST build compiled with
AddressSanitizer (see tests/emu/asan_case.py).  Stands in for compute-sanitizer memcheck, which is closed
on the GPU pool."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_kernels_are_asan_clean(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    objs = []
    for s in ("orbfe_api", "orbfe_match", "orbfe_bow", "orbfe_frame"):
        o = str(tmp_path / (s + ".o"))
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-fsanitize=address", "-fno-omit-frame-pointer",
                               "-ffp-contract=off", "-fno-strict-aliasing", "-DORBFE_EMU", "-include",
                               os.path.join(HERE, "emu", "cuda_emu.h"), "-x", "c++", "-c",
                               os.path.join(ROOT, "slam_framework_b200", "csrc", s + ".cu"), "-o", o])
        objs.append(o)
    lib = str(tmp_path / "liborbfe_emu_asan_TESTONLY.so")
    subprocess.check_call(["g++", "-shared", "-fsanitize=address", "-o", lib] + objs)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "asan_case.py"), lib, ROOT], env=env, capture_output=True,
                       text=True, timeout=900)
    assert "ERROR: AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0 and "ASAN-RUN-OK" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
