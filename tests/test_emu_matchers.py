"""Matcher kernel LOGIC on CPU (emulated TEST build, see test_emu_parity.py) against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
import parity_common as P
from emu import build_emu
from slam_framework_b200 import orbfe, synth


@pytest.fixture(scope="module")
def emu():
    return orbfe.load(build_emu.build(), _test_emulation=True)


def oracle_extract(img, nfeatures):
    return O.Extractor(nfeatures).extract(img)


def test_descriptor_distance(emu):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10:20] = ~a[10:20]
    d = orbfe.DescriptorDistance(a, b, lib=emu)
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(1))
    assert d[:10].max() == 0 and d[10:20].min() == 256


def test_search_for_initialization_small(emu):
    a, b = synth.shifted_frame(3, 200, 640, dx=8, dy=4)
    assert P.check_search_for_initialization(emu, a, b, oracle_extract, nfeatures=1500) > 20


def test_search_by_projection_mappoints(emu):
    img = synth.frame(240, 800, seed=2)
    kps, desc = oracle_extract(img, 1500)
    scale = O.Extractor(1500).tables()["scale"]
    rng = np.random.default_rng(1)
    ur = np.where(rng.uniform(0, 1, len(kps)) < 0.6, kps["x"] - rng.uniform(1, 60, len(kps)), -1).astype(np.float32)
    assert P.check_search_by_projection_mappoints(emu, kps, desc, scale, 800, 240, 3000, seed=5, u_right=ur) > 200
    assert P.check_search_by_projection_mappoints(emu, kps, desc, scale, 800, 240, 2000, seed=6, th=3) > 100


def test_search_by_projection_lastframe(emu):
    img = synth.frame(240, 800, seed=4)
    kps, desc = oracle_extract(img, 1200)
    scale = O.Extractor(1200).tables()["scale"]
    rng = np.random.default_rng(2)
    ur = np.where(rng.uniform(0, 1, len(kps)) < 0.5, kps["x"] - rng.uniform(1, 60, len(kps)), -1).astype(np.float32)
    assert P.check_search_by_projection_lastframe(emu, kps, desc, scale, 800, 240, seed=7, u_right=ur) > 100


def test_empty_frames(emu):
    scale = O.Extractor().tables()["scale"]
    F = orbfe.Frame(np.zeros(0, orbfe.KP_DTYPE), np.zeros((0, 32), np.uint8), scale, (0, 100, 0, 100), lib=emu)
    assert len(F.GetFeaturesInArea(10, 10, 5)) == 0
    n, m, _ = orbfe.OrbMatcher(0.9).SearchForInitialization(F, F, np.zeros((0, 2), np.float32), 100)
    assert n == 0 and len(m) == 0


def test_search_by_bow(emu):
    a, b = synth.shifted_frame(5, 200, 640, dx=6, dy=3)
    ka, da = oracle_extract(a, 1200)
    kb, db = oracle_extract(b, 1200)
    scale = O.Extractor(1200).tables()["scale"]
    assert P.check_search_by_bow(emu, kb, db, ka, da, scale, 640, 200, seed=3) > 60
    assert P.check_search_by_bow(emu, ka, da, ka, da, scale, 640, 200, seed=4, nnratio=0.75) > 300  # frame vs itself


# ---- N1: the remaining OrbMatcher searches -----------------------------------------------------------------
@pytest.fixture(scope="module")
def two_frames():
    a, b = synth.shifted_frame(7, 200, 640, dx=6, dy=0)
    ka, da = oracle_extract(a, 1200)
    kb, db = oracle_extract(b, 1200)
    return ka, da, kb, db, O.Extractor(1200).tables()["scale"]


def test_search_by_projection_sim3(emu, two_frames):
    ka, da, _, _, scale = two_frames
    assert P.check_search_by_projection_sim3(emu, ka, da, scale, 640, 200, 2500, seed=11) > 150


def test_search_by_projection_keyframe(emu, two_frames):
    ka, da, _, _, scale = two_frames
    assert P.check_search_by_projection_keyframe(emu, ka, da, scale, 640, 200, 1500, seed=12) > 400


def test_fuse(emu, two_frames):
    ka, da, _, _, scale = two_frames
    rng = np.random.default_rng(5)
    ur = np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - rng.uniform(1, 60, len(ka)), -1).astype(np.float32)
    assert P.check_fuse(emu, ka, da, scale, 640, 200, 2500, seed=13, u_right=ur) > 300


def test_search_by_sim3(emu, two_frames):
    ka, da, kb, db, scale = two_frames
    assert P.check_search_by_sim3(emu, ka, da, kb, db, scale, 640, 200, seed=14, shift=(6.0, 0.0)) > 100


def test_search_by_bow_keyframes(emu, two_frames):
    ka, da, kb, db, scale = two_frames
    assert P.check_search_by_bow_keyframes(emu, ka, da, kb, db, scale, 640, 200, seed=15) > 60


def test_search_for_triangulation(emu, two_frames):
    ka, da, kb, db, scale = two_frames
    assert P.check_search_for_triangulation(emu, ka, da, kb, db, scale, 640, 200, seed=16) > 100


# ---- N3: Frame::ComputeBoW -----------------------------------------------------------------------------------
def test_bow_transform(emu, two_frames, tmp_path):
    ka, da, _, _, _ = two_frames
    assert P.check_bow_transform(emu, da, seed=21, k=10, L=3, tmp_path=tmp_path) > 1000
    assert P.check_bow_transform(emu, da[:300], seed=22, k=4, L=5) > 300


def test_bow_transform_degenerate(emu):
    rng = np.random.default_rng(3)
    arrays = P.synth_vocabulary(rng, 3, 2)
    V = orbfe.OrbVocabulary(3, 2, 0, 0, *arrays, lib=emu)
    out = V.transform(np.zeros((0, 32), np.uint8))
    assert len(out["bow"][0]) == 0 and len(out["fv"][0]) == 0
    one = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    P.assert_bow_equal(V.transform(one, 1), O.Vocabulary(3, 2, 0, 0, *arrays).transform(one, 1))
    # all words stopped: nothing survives
    arrays0 = (arrays[0], arrays[1], arrays[2], np.zeros_like(arrays[3]))
    out = orbfe.OrbVocabulary(3, 2, 0, 0, *arrays0, lib=emu).transform(rng.integers(0, 256, (50, 32), dtype=np.uint8))
    assert len(out["bow"][0]) == 0 and len(out["fv"][0]) == 0


# ---- N2: the Frame tail --------------------------------------------------------------------------------------------
def test_undistort_keypoints(emu, two_frames):
    ka = two_frames[0]
    assert P.check_undistort_keypoints(emu, ka, seed=31) > 500
    assert len(orbfe.UndistortKeyPoints(ka[:0], 700.0, 700.0, 600.0, 180.0, [0.1, 0, 0, 0], lib=emu)) == 0


def test_is_in_frustum(emu):
    assert P.check_is_in_frustum(emu, 20000, seed=32) > 2000
    assert P.check_is_in_frustum(emu, 333, seed=33) > 30


def test_glibc_logf_restatement_matches_libm(emu):
    assert P.check_logf(emu, 200000, seed=5) > 50000


def test_search_for_initialization_acceptor_overflow_falls_back_to_serial(emu):
    """20 queries take the same keypoint one after the other at ever smaller distances (each accept lowers vMatchedDistance):
    more acceptors than the parallel resolve keeps per keypoint, so it must hand over to the serial kernel; plus a longer
    stealing chain inside the slot budget"""
    scale = O.Extractor().tables()["scale"]
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    for n1 in (20, 6):
        k2 = np.zeros(3, orbfe.KP_DTYPE)
        k2["x"], k2["y"], k2["angle"] = [100, 140, 300], [80, 90, 200], [10, 20, 30]
        d2 = np.stack([base, ~base, rng.integers(0, 256, 32, dtype=np.uint8)])
        k1 = np.zeros(n1, orbfe.KP_DTYPE)
        k1["x"], k1["y"] = 102 + np.arange(n1) % 3, 81
        k1["angle"] = (np.arange(n1) * 7) % 360
        d1 = np.repeat(base[None, :], n1, 0)
        for i in range(n1):            # distance to `base` = n1 - i: strictly decreasing along the query order
            for b in range(n1 - i):
                d1[i, b >> 3] ^= np.uint8(1 << (b & 7))
        bounds = (0.0, 640.0, 0.0, 480.0)
        F1, F2 = orbfe.Frame(k1, d1, scale, bounds, lib=emu), orbfe.Frame(k2, d2, scale, bounds, lib=emu)
        O1, O2 = O.Frame(k1, d1, scale, bounds), O.Frame(k2, d2, scale, bounds)
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        for ori in (True, False):
            n, m12, pm = orbfe.OrbMatcher(0.9, ori).SearchForInitialization(F1, F2, prev, 100)
            on, om12, opm = O.search_for_initialization(O1, O2, prev, 100, 0.9, ori)
            assert n == on and np.array_equal(m12, om12) and np.array_equal(pm, opm)
            if not ori:   # the last (closest) query owns the keypoint; with the rotation check its lone bin survives or not as in the oracle
                assert n == 1 and m12[n1 - 1] == 0 and (m12[: n1 - 1] == -1).all()


def test_search_local_points_device_resident(emu, two_frames):
    ka, da, _, _, scale = two_frames
    rng = np.random.default_rng(41)
    ur = np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - rng.uniform(1, 60, len(ka)), -1).astype(np.float32)
    # the synthetic camera of the check is KITTI's; a 640x200 frame sees the upper-left part of its field of view
    assert P.check_search_local_points(emu, ka, da, scale, 640, 200, seed=42, u_right=ur) > 200
    assert P.check_search_local_points(emu, ka, da, scale, 640, 200, seed=43, th=3, n_extra=500) > 200


def test_empty_inputs_everywhere(emu, two_frames):
    ka, da, _, _, scale = two_frames
    P.check_empty_inputs(emu, ka, da, scale)


def test_frame_from_extractor_device_resident(emu):
    l, r = synth.stereo_pair(188, 620, seed=6)
    assert P.check_frame_from_extractor(emu, l, r, nfeatures=800, seed=51) > 100


def test_fuzz_slice_matchers_on_tie_heavy_keypoints():
    """tests/fuzz_matchers.py (every search routine on keypoints from checkerboards / rectangles / gratings: identical
    descriptors everywhere, so the tie rules decide) -- seeds 0..5 on the emulated build"""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "fuzz_matchers.py"), "emu", "6"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "failures 0" in r.stdout


def test_frame_handles_are_recycled_and_reset(emu):
    """orbfe_frame_destroy parks a handle, the next orbfe_frame_create takes it over: a recycled handle must behave like a new one
    (other keypoint count, bounds, level table), orbfe_frame_pool_trim frees the idle ones, and ORBFE_NO_HANDLE_POOL-free
    operation is what every other test already runs on."""
    emu.orbfe_frame_pool_trim()
    big = synth.frame(240, 800, seed=11)
    small = synth.frame(120, 400, seed=12)
    kb, db = oracle_extract(big, 1500)
    ks, ds = O.Extractor(400, 1.5, 5, 20, 7).extract(small)
    sb, ss = O.Extractor(1500).tables()["scale"], O.Extractor(400, 1.5, 5, 20, 7).tables()["scale"]
    for rounds in range(2):   # big -> small -> big on recycled handles
        assert P.check_search_by_projection_mappoints(emu, kb, db, sb, 800, 240, 1200, seed=3) > 50
        assert P.check_search_by_projection_mappoints(emu, ks, ds, ss, 400, 120, 300, seed=4) >= 0
        assert P.check_search_by_projection_lastframe(emu, ks, ds, ss, 400, 120, seed=5) >= 0
    import gc
    gc.collect()                                   # the Frame wrappers above are gone: their handles sit in the pool
    assert emu.orbfe_frame_pool_trim() >= 1
    assert emu.orbfe_frame_pool_trim() == 0
    assert P.check_search_by_projection_mappoints(emu, kb, db, sb, 800, 240, 1200, seed=3) > 50   # fresh handles again


def test_pinned_host_buffers(emu):
    """orbfe_pinned_alloc through the numpy wrapper: frames uploaded from it and results downloaded into it are the same bytes"""
    l, r = synth.stereo_pair(120, 400, seed=21)
    frames = orbfe.pinned_empty((2, 120, 400), np.uint8, lib=emu)
    frames[0], frames[1] = l, r
    ex = orbfe.ORBextractor(400, lib=emu, max_images=2)
    ex.upload([frames[0], frames[1]]); ex.run(2)
    cap = ex.max_keypoints()
    buf = dict(kps=orbfe.pinned_empty((2, cap), orbfe.KP_DTYPE, lib=emu), desc=orbfe.pinned_empty((2, cap, 32), np.uint8, lib=emu),
               n=orbfe.pinned_empty((2,), np.int32, lib=emu), cap=cap, ur=None, depth=None)
    ex.download_async(2, buf); ex.sync()
    k0, d0 = oracle_extract(l, 400)
    n0 = int(buf["n"][0])
    assert n0 == len(k0) and np.array_equal(buf["desc"][0, :n0], d0) and np.array_equal(buf["kps"]["x"][0, :n0], k0["x"])
    ex.close()
    orbfe.release_pinned()
