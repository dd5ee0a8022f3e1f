"""The header-only C++ drop-in shim (include/orbfe_shim.hpp) compiles against the OpenCV stand-ins
(include/cv_compat.h; real OpenCV C++ is absent here), links liborbfe.so, and - on the GPU box - produces
the oracle's keypoints, descriptors and stereo matches when driven like the reference's Frame ctor."""
import os
import subprocess

import numpy as np
import pytest

from slam_framework_b200 import build as B
from slam_framework_b200 import orbfe, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "shim", "shim_demo")


def build_demo():
    lib = B.build()
    src = os.path.join(ROOT, "tests", "shim", "shim_demo.cpp")
    deps = [src, os.path.join(ROOT, "include", "orbfe_shim.hpp"), os.path.join(ROOT, "include", "cv_compat.h"), lib]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               lib, "-Wl,-rpath," + os.path.dirname(lib), "-pthread"])
    return EXE


def test_shim_compiles_as_cpp11_and_links():
    out = subprocess.check_output([build_demo(), "--link-check"], text=True)
    assert out.startswith("orbfe 0.1 (sm_100a)")


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path):
    import oracle_lib as O
    l, r = synth.stereo_pair(seed=77)
    l.tofile(tmp_path / "l.raw"); r.tofile(tmp_path / "r.raw")
    out = subprocess.check_output([build_demo(), "1241", "376", str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                                   str(tmp_path / "o")], text=True)
    nl, nr, pw, ph = map(int, out.split())
    oL, oR = O.Extractor(), O.Extractor()
    okl, odl = oL.extract(l); okr, odr = oR.extract(r)
    kl = np.fromfile(tmp_path / "o.kl", orbfe.KP_DTYPE); kr = np.fromfile(tmp_path / "o.kr", orbfe.KP_DTYPE)
    assert nl == len(okl) and nr == len(okr)
    assert np.array_equal(kl, okl) and np.array_equal(kr, okr)
    assert np.array_equal(np.fromfile(tmp_path / "o.dl", np.uint8).reshape(-1, 32), odl)
    assert np.array_equal(np.fromfile(tmp_path / "o.dr", np.uint8).reshape(-1, 32), odr)
    baseline = float(np.float32(386.1448) / np.float32(718.856))  # the demo divides in float
    on, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, 386.1448, baseline)
    assert np.array_equal(np.fromfile(tmp_path / "o.ur", np.float32), our)
    assert np.array_equal(np.fromfile(tmp_path / "o.depth", np.float32), odp)
    assert np.array_equal(np.fromfile(tmp_path / "o.pyr3", np.uint8).reshape(ph, pw), oL.pyramid_level(3))


@pytest.mark.gpu
def test_shim_n1_adapters_find_a_frame_in_itself(tmp_path):
    """every N1 adapter of orbfe_shim.hpp (SearchByBoW x2, SearchForTriangulation, SearchByProjection x2, SearchBySim3,
    Fuse x2) instantiated over stand-in KeyFrame / MapPoint / Frame types and run on the GPU: a frame searched in itself
    must match (nearly) every keypoint to itself"""
    img = synth.frame(240, 800, seed=5)
    img.tofile(tmp_path / "i.raw")
    out = subprocess.check_output([build_demo(), "--n1", "800", "240", str(tmp_path / "i.raw")], text=True)
    n, bow_f, bow_kf, tri, proj_kf, proj_f, sim3, fuse1, fuse2, moved, visible, at_pixel = map(int, out.split())
    assert n > 500
    for got in (bow_f, bow_kf, tri, proj_kf, proj_f, sim3, fuse1, fuse2):
        assert got > 0.5 * n, out
    # N2 adapters: a real distortion model moves (nearly) every keypoint; every back-projected point is visible at its pixel
    assert moved > 0.9 * n and visible == n and at_pixel == n, out


@pytest.mark.gpu
def test_shim_vocabulary_transform(tmp_path):
    """orbfe::OrbVocabulary (loadFromTextFile + transform into std::map containers) == the oracle"""
    import oracle_lib as O
    import parity_common as P
    img = synth.frame(240, 800, seed=6)
    img.tofile(tmp_path / "i.raw")
    kps, desc = O.Extractor(1000).extract(img)
    rng = np.random.default_rng(9)
    arrays = P.synth_vocabulary(rng, 10, 3, seed_desc=desc[0].copy())
    P.write_vocabulary_text(str(tmp_path / "voc.txt"), 10, 3, 0, 0, arrays)
    out = subprocess.check_output([build_demo(), "--bow", "800", "240", str(tmp_path / "i.raw"), str(tmp_path / "voc.txt")], text=True)
    nb, nf, nfeat, total = out.split()
    ref = O.Vocabulary(10, 3, 0, 0, *arrays).transform(desc, 1)
    assert int(nb) == len(ref["bow"][0]) and int(nf) == len(ref["fv"][0]) and int(nfeat) == len(ref["fv"][2])
    s = 0.0
    for v in ref["bow"][1]:
        s += float(v)
    assert float(total) == s
