"""The drop-in as compiled code: dropin/ (class ORBextractor, class OrbMatcher as declared in the reference's own orb_matcher.h,
Frame::ComputeStereoMatches) linked with the reference's real Frame / KeyFrame / MapPoint / Map / DBoW2 sources.  Every test runs
one scenario twice on the reference's own objects -- through oracle/_ref (the reference's bodies) and through the drop-in library
(the GPU bodies) -- and demands identical results: keypoints, descriptors, stereo coordinates, map-point assignments, match
vectors, fused / replaced points and return counts, for all eleven OrbMatcher routines and the stereo Frame constructor, with
rotated poses and a Sim3 scale != 1 as well.

`-m gpu`: the library linked against liborbfe.so.  `-m "not gpu"`: the same tests against the emulated kernel build (slow)."""
import numpy as np
import pytest

import dropin_lib
import parity_common as P
import reference_lib as R
from slam_framework_b200 import synth

CAM = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448)
f32 = np.float32


def _rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (Rz @ Ry @ Rx).astype(np.float32)


ROTS = [None, _rot(0.02, -0.015, 0.01), _rot(-0.03, 0.025, -0.04)]


def _kinds():
    out = []
    if R.available() and dropin_lib.available("emu"):
        out.append(pytest.param("emu", id="emulated"))
    if dropin_lib.available("gpu"):
        out.append(pytest.param("gpu", id="gpu", marks=pytest.mark.gpu))
    return out


@pytest.fixture(scope="module", params=_kinds())
def D(request):
    if not R.available():
        pytest.skip("oracle/_ref is not available")
    if request.param == "gpu":
        from slam_framework_b200 import orbfe
        if orbfe.load().orbfe_device_count() < 1:
            pytest.skip("no CUDA device")
    return dropin_lib.load(request.param)


def both(D, fn):
    """fn(M) evaluated with M = the reference binding and M = the drop-in binding"""
    return fn(R), fn(D)


def same(a, b, what):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), what
        for k in a:
            same(a[k], b[k], f"{what}[{k}]")
    elif isinstance(a, (tuple, list)):
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{what}[{i}]")
    elif isinstance(a, np.ndarray):
        if a.dtype.names:
            for f in a.dtype.names:
                assert np.array_equal(a[f], b[f]), f"{what}.{f} differs ({(a[f] != b[f]).sum()} entries)"
        else:
            assert a.shape == b.shape and np.array_equal(a, b), f"{what} differs ({(np.asarray(a) != np.asarray(b)).sum()} entries)"
    else:
        assert a == b, (what, a, b)


def frame_state(F):
    return dict(n=F.n, kps=F.kps, kps_un=F.kps_un, desc=F.desc, u_right=F.u_right, depth=F.depth, bounds=F.bounds, scale=F.scale)


def test_extractor_compute(D):
    """ORBextractor::Compute (orb_extractor.cpp:985-1049) behind the reference's signature"""
    for img, nf, params in ((synth.frame(200, 640, seed=3), 800, (1.2, 8, 20, 7)), (synth.frame(160, 500, seed=4), 500, (1.5, 4, 20, 7))):
        (k, d), (dk, dd) = both(D, lambda M: M.extract(img, nf, *params))
        assert len(k) > 100
        same(k, dk, "keypoints")
        same(d, dd, "descriptors")
    same(*both(D, lambda M: M.tables(1000, 1.2, 8)), "scale tables")


@pytest.mark.parametrize("seed,shape,nf", [(3, (376, 1241), 2000), (5, (240, 800), 1000)])
def test_stereo_frame_constructor(D, seed, shape, nf):
    """Frame's stereo constructor (frame.cpp:61-111): two ORBextractor::Compute on two std::threads, UndistortKeyPoints (the
    reference's), Frame::ComputeStereoMatches (frame.cpp:406-577, the drop-in's), AssignFeaturesToGrid (the reference's)"""
    l, r = synth.stereo_pair(*shape, seed=seed)
    FR, FD = both(D, lambda M: M.Frame(l, r, nfeatures=nf))
    same(frame_state(FR), frame_state(FD), "stereo frame")
    assert (FR.u_right >= 0).sum() > 100
    rng = np.random.default_rng(seed)
    for _ in range(40):   # the grid the matchers walk is the reference's own in both libraries
        x, y, rad = rng.uniform(0, shape[1]), rng.uniform(0, shape[0]), float(rng.choice([8.0, 30.0]))
        assert np.array_equal(FR.features_in_area(x, y, rad), FD.features_in_area(x, y, rad))


def test_search_for_initialization(D):
    a, b = synth.shifted_frame(21, 240, 800, dx=8, dy=4)
    for window, ratio, ori in ((100, 0.9, True), (40, 0.8, False)):
        def run(M):
            F1, F2 = M.Frame(a, nfeatures=2500), M.Frame(b, nfeatures=2500)
            prev = np.stack([F1.kps_un["x"], F1.kps_un["y"]], 1).astype(np.float32)
            return M.search_for_initialization(F1, F2, prev, window, ratio, ori)
        r, d = both(D, run)
        assert r[0] > 50
        same(r, d, "SearchForInitialization")


def test_search_by_projection_mappoints(D):
    l, r = synth.stereo_pair(240, 800, seed=31)
    for seed, n_mp, th in ((5, 4000, 1), (6, 2000, 3)):
        def run(M):
            F = M.Frame(l, r, nfeatures=1500)
            mp = P.synth_map_points(F.kps_un, F.desc, np.random.default_rng(seed), n_mp, F.u_right)
            args = (mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"], mp["desc"], mp["has_obs"], mp["occupied"])
            return M.search_by_projection_mappoints(F, *args, th, 0.8)
        r_, d_ = both(D, run)
        assert r_[0] > 200
        same(r_, d_, "SearchByProjection(Frame, MapPoints)")


@pytest.mark.parametrize("rot", range(len(ROTS)))
def test_search_by_projection_lastframe(D, rot):
    l, r = synth.stereo_pair(240, 800, seed=33)

    def run(M):
        M.set_test_transform(ROTS[rot])
        try:
            cur, last = M.Frame(l, r, nfeatures=1500), M.Frame(l, r, nfeatures=1500)
            n = cur.n
            rng = np.random.default_rng(9)
            fx, cx, cy = (f32(CAM[k]) for k in ("fx", "cx", "cy"))
            z = rng.uniform(3, 60, n).astype(f32)
            z[:7] = -5.0
            xw = ((last.kps_un["x"] + rng.uniform(-5, 5, n).astype(f32) - cx) / fx * z).astype(f32)
            yw = ((last.kps_un["y"] + rng.uniform(-5, 5, n).astype(f32) - cy) / fx * z).astype(f32)
            xw[7:12] *= 40
            world = np.stack([xw, yw, z], 1).astype(f32)
            d = last.desc.copy()
            d[:, 0] ^= rng.integers(0, 256, n).astype(np.uint8)
            valid = (rng.uniform(0, 1, n) < 0.9).astype(np.uint8)
            has_obs = (rng.uniform(0, 1, n) < 0.7).astype(np.uint8)
            occupied = (rng.uniform(0, 1, n) < 0.05).astype(np.uint8)
            base = float(cur.misc[0])
            out = []
            for last_t, mono in (((0, 0, 0), False), ((0, 0, 2 * base), False), ((0, 0, -2 * base), False), ((0, 0, 2 * base), True)):
                for ori in (True, False):
                    out.append(M.search_by_projection_lastframe(cur, last, valid, world, d, has_obs, np.array(last_t, f32), occupied,
                                                                7.0, mono, ori))
            return out
        finally:
            M.set_test_transform(None)
    r_, d_ = both(D, run)
    assert max(x[0] for x in r_) > 200
    same(r_, d_, "SearchByProjection(Cur, Last)")


@pytest.fixture(scope="module")
def voc_path(tmp_path_factory):
    rng = np.random.default_rng(12)
    arrays = P.synth_vocabulary_uniform(rng, 6, 5, seed_desc=None)
    path = tmp_path_factory.mktemp("voc") / "voc_6_5.txt"
    P.write_vocabulary_text(str(path), 6, 5, 0, 0, arrays)
    return path


def bow_frames(M, voc_path, a, b):
    V = M.Vocabulary(voc_path)
    M.set_vocabulary(V)
    try:
        return V, M.Frame(a, a, nfeatures=1500), M.Frame(b, b, nfeatures=1500)
    finally:
        M.set_vocabulary(None)


def test_search_by_bow_and_triangulation(D, voc_path):
    """SearchByBoW x2 (orb_matcher.cpp:133-262, 499-632) and SearchForTriangulation (:634-802)"""
    a, b = synth.shifted_frame(43, 240, 800, dx=6, dy=0)

    def run(M):
        V, FA, FB = bow_frames(M, voc_path, a, b)
        rng = np.random.default_rng(3)
        va = (rng.uniform(0, 1, FA.n) < 0.7).astype(np.uint8)
        vb = (rng.uniform(0, 1, FB.n) < 0.7).astype(np.uint8)
        bad_a = ((rng.uniform(0, 1, FA.n) < 0.05) & (va == 1)).astype(np.uint8)
        bad_b = ((rng.uniform(0, 1, FB.n) < 0.05) & (vb == 1)).astype(np.uint8)
        KA, KB = M.KeyFrame(FA, va, bad_a), M.KeyFrame(FB, vb, bad_b)
        out = []
        for ratio, ori in ((0.7, True), (0.9, False)):
            out.append(M.search_by_bow_kf_f(KA, FB, ratio, ori))
            out.append(M.search_by_bow_kf_kf(KA, KB, ratio, ori))
        has_a = (rng.uniform(0, 1, FA.n) < 0.2).astype(np.uint8)
        has_b = (rng.uniform(0, 1, FB.n) < 0.2).astype(np.uint8)
        Fa = np.array([[0, 0, 0], [0, 0, 6.0], [0, -6.0, 0]], f32)
        for t2 in ((-0.5, 0.0, 0.05), (0.3, 0.02, 1.0)):
            KA2, KB2 = M.KeyFrame(FA, has_a), M.KeyFrame(FB, has_b, translation=t2)
            for F12 in (Fa, (Fa + np.random.default_rng(4).normal(0, 2e-4, (3, 3))).astype(f32)):
                for only_stereo in (False, True):
                    out.append(M.search_for_triangulation(KA2, KB2, F12, only_stereo, True))
        return out
    r_, d_ = both(D, run)
    assert r_[0][0] > 50 and max(x[0] for x in r_[4:]) > 50
    same(r_, d_, "SearchByBoW / SearchForTriangulation")


def scene_points(F, rng, n, jitter=3.0):
    fx, cx, cy = (f32(CAM[k]) for k in ("fx", "cx", "cy"))
    idx = rng.integers(0, F.n, n).astype(np.int32)
    z = rng.uniform(4, 50, n).astype(f32)
    z[: n // 40] = rng.uniform(-8, -1, n // 40).astype(f32)
    x = ((F.kps_un["x"][idx] + rng.uniform(-jitter, jitter, n).astype(f32) - cx) / fx * z).astype(f32)
    y = ((F.kps_un["y"][idx] + rng.uniform(-jitter, jitter, n).astype(f32) - cy) / fx * z).astype(f32)
    x[n // 40: n // 20] *= f32(30)
    world = np.stack([x, y, z], 1).astype(f32)
    d = F.desc[idx].copy()
    flip = rng.integers(0, 256, (n, 24))
    for i in range(n):
        for b in flip[i, : rng.integers(0, 24)]:
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return idx, world, d


@pytest.mark.parametrize("rot,scale", [(0, 1.0), (1, 1.0), (2, 1.07)])
def test_keyframe_projection_searches(D, rot, scale):
    """SearchByProjection(KeyFrame*, Scw, ...) (:384-497), Fuse(KeyFrame*, Scw, ...) (:956-1079), Fuse(KeyFrame*, MapPoints, th)
    (:804-954) with the rotation / Sim3 scale of the case"""
    l, r = synth.stereo_pair(240, 800, seed=37)

    def run(M):
        M.set_test_transform(ROTS[rot], scale)
        try:
            F = M.Frame(l, r, nfeatures=1500)
            rng = np.random.default_rng(14)
            out = []
            for t, th in (((0.0, 0.0, 0.0), 10), ((0.3, -0.1, 1.5), 10), ((-0.2, 0.05, -2.0), 4)):
                idx, world, d = scene_points(F, rng, 2000)
                bad = (rng.uniform(0, 1, len(idx)) < 0.05).astype(np.uint8)
                matched_in = (rng.uniform(0, 1, F.n) < 0.2).astype(np.uint8)
                n, m, _ = M.search_by_projection_sim3(M.KeyFrame(F), world, idx, d, bad, matched_in, t, th)
                out.append((n, m))
                n, best, _ = M.fuse_sim3(M.KeyFrame(F), world, idx, d, bad, t, float(th))
                out.append((n, best))
                n, best, _ = M.fuse(M.KeyFrame(F), world, idx, d, bad, t, 3.0)
                out.append((n, best))
            return out
        finally:
            M.set_test_transform(None)
    r_, d_ = both(D, run)
    assert max(x[0] for x in r_) > 100
    same(r_, d_, "KeyFrame projection searches")


@pytest.mark.parametrize("rot,scale", [(0, 1.0), (1, 0.95)])
def test_relocalisation_search_and_sim3(D, rot, scale):
    """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (:1455-1582) and SearchBySim3 (:1081-1310)"""
    a, b = synth.shifted_frame(45, 240, 800, dx=5, dy=2)

    def run(M):
        M.set_test_transform(ROTS[rot], scale)
        try:
            FA, FB = M.Frame(a, a, nfeatures=1500), M.Frame(b, b, nfeatures=1500)
            rng = np.random.default_rng(15)
            fx, cx, cy = (f32(CAM[k]) for k in ("fx", "cx", "cy"))

            def own_points(F):
                n = F.n
                z = rng.uniform(4, 50, n).astype(f32)
                world = np.stack([((F.kps_un["x"] - cx) / fx * z).astype(f32), ((F.kps_un["y"] - cy) / fx * z).astype(f32), z], 1).astype(f32)
                d = F.desc.copy()
                d[:, 5] ^= rng.integers(0, 8, n).astype(np.uint8)
                valid = (rng.uniform(0, 1, n) < 0.8).astype(np.uint8)
                return valid, world, d
            va, wa, da = own_points(FA)
            KA = M.KeyFrameAt(FA, va, wa, da)
            out = []
            for cur_t, th, orb_dist in (((0.0, 0.0, 0.0), 10.0, 100), ((0.05, 0.02, 0.8), 15.0, 64), ((0.0, 0.0, -30.0), 10.0, 100)):
                found = ((rng.uniform(0, 1, FA.n) < 0.1) & (va == 1)).astype(np.uint8)
                occupied = (rng.uniform(0, 1, FB.n) < 0.1).astype(np.uint8)
                for ori in (True, False):
                    out.append(M.search_by_projection_keyframe(FB, KA, found, cur_t, occupied, th, orb_dist, ori))
            vb, wb, db = own_points(FB)
            KB = M.KeyFrameAt(FB, vb, wb, db)
            for t12, th in (((0.0, 0.0, 0.0), 7.5), ((0.02, -0.01, 0.3), 7.5)):
                pre = np.full(FA.n, -1, np.int32)
                cand = np.where((va == 1) & (rng.uniform(0, 1, FA.n) < 0.05))[0]
                pre[cand] = rng.choice(np.where(vb == 1)[0], len(cand), replace=False)
                out.append(M.search_by_sim3(KA, KB, np.asarray(t12, f32), th, pre))
            return out
        finally:
            M.set_test_transform(None)
    r_, d_ = both(D, run)
    assert max(x[0] for x in r_) > 100
    same(r_, d_, "relocalisation search / SearchBySim3")
