"""Shared parity checks: the C-ABI library (product build on the GPU, or the emulated TEST build on
CPU) against the CPU oracle on the same seeded inputs.  Bit-exact for keypoints (coordinates,
octave, response, size), descriptors, match indices; angles / sub-pixel disparities are compared
bit-exact too (the stated tolerance of BASELINE.json is 1e-4; we meet 0)."""
import numpy as np

import oracle_lib as O
from slam_framework_b200 import orbfe

KITTI = dict(bf=386.1448, fx=718.856)
KP_FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")
ANGLE_TOL = 1e-4


def assert_kps_equal(got, ref, what=""):
    assert len(got) == len(ref), f"{what}: {len(got)} keypoints vs oracle {len(ref)}"
    for f in KP_FIELDS:
        if f == "angle":
            assert np.allclose(got[f], ref[f], rtol=0, atol=ANGLE_TOL), f"{what}: angle beyond {ANGLE_TOL}"
        assert np.array_equal(got[f], ref[f]), f"{what}: field {f} differs"


def check_extract(lib, img, nfeatures=2000, stages=True, params=(1.2, 8, 20, 7)):
    ex = orbfe.ORBextractor(nfeatures, *params, lib=lib)
    oe = O.Extractor(nfeatures, *params)
    kps, desc = ex.Compute(img)
    okps, odesc = oe.extract(img)
    if stages:
        for l in range(params[1]):
            assert np.array_equal(ex.pyramid_level(l), oe.pyramid_level(l)), f"pyramid level {l}"
            c, oc = ex.debug_candidates(l), oe.candidates(l)
            assert len(c) == len(oc), f"FAST candidates level {l}: {len(c)} vs {len(oc)}"
            for f in ("x", "y", "response"):
                assert np.array_equal(c[f], oc[f]), f"FAST candidates level {l} field {f}"
            k, ok = ex.debug_level_keypoints(l), oe.level_keypoints(l)
            assert len(k) == len(ok), f"quad-tree level {l}: {len(k)} vs {len(ok)}"
            assert np.array_equal(k["x"] + 16, ok["x"]) and np.array_equal(k["y"] + 16, ok["y"]), f"quad-tree level {l}"
            if len(ok):
                assert np.array_equal(ex.debug_blurred(l), oe.blurred(l)), f"blur level {l}"
    assert_kps_equal(kps, okps, "extract")
    assert np.array_equal(desc, odesc), f"descriptors differ in {(desc != odesc).any(1).sum()} rows"
    ex.close()
    return kps, desc


def check_stereo(lib, left, right, nfeatures=2000, bf=KITTI["bf"], fx=KITTI["fx"], params=(1.2, 8, 20, 7)):
    eL, eR = orbfe.ORBextractor(nfeatures, *params, lib=lib), orbfe.ORBextractor(nfeatures, *params, lib=lib)
    oL, oR = O.Extractor(nfeatures, *params), O.Extractor(nfeatures, *params)
    kl, dl = eL.Compute(left)
    kr, dr = eR.Compute(right)
    okl, odl = oL.extract(left)
    okr, odr = oR.extract(right)
    assert_kps_equal(kl, okl, "left")
    assert_kps_equal(kr, okr, "right")
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    n, ur, dp = orbfe.ComputeStereoMatches(eL, eR, kl, dl, kr, dr, bf, bf / fx)
    on, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, bf / fx)
    assert n == on, f"stereo matches {n} vs oracle {on}"
    assert np.array_equal(ur >= 0, our >= 0), "matched set differs"
    assert np.allclose(ur, our, rtol=0, atol=1e-4) and np.allclose(dp, odp, rtol=1e-6, atol=1e-4)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp), "sub-pixel results not bit-identical"
    eL.close(); eR.close()
    return n


def check_batch_stereo(lib, pairs, nfeatures=2000, bf=KITTI["bf"], fx=KITTI["fx"]):
    """device-resident batch path == per-frame oracle, slot by slot."""
    n = 2 * len(pairs)
    ex = orbfe.ORBextractor(nfeatures, lib=lib, max_images=n)
    imgs = [im for p in pairs for im in p]
    ex.upload(imgs)
    ex.run(n)
    ex.run_stereo(len(pairs), bf, bf / fx)
    buf = ex.download(n, ex.make_buffers(n, stereo=True))
    total = 0
    for p, (l, r) in enumerate(pairs):
        oL, oR = O.Extractor(nfeatures), O.Extractor(nfeatures)
        okl, odl = oL.extract(l)
        okr, odr = oR.extract(r)
        nl, nr = buf["n"][2 * p], buf["n"][2 * p + 1]
        assert_kps_equal(buf["kps"][2 * p, :nl], okl, f"pair {p} left")
        assert_kps_equal(buf["kps"][2 * p + 1, :nr], okr, f"pair {p} right")
        assert np.array_equal(buf["desc"][2 * p, :nl], odl) and np.array_equal(buf["desc"][2 * p + 1, :nr], odr)
        on, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, bf / fx)
        assert np.array_equal(buf["ur"][2 * p, :nl], our), f"pair {p} uR"
        assert np.array_equal(buf["depth"][2 * p, :nl], odp), f"pair {p} depth"
        total += on
    ex.close()
    return total


# ---- matchers (BASELINE configs 4 and 5, per-frame tracking) --------------------------------------
def make_frames(kps, desc, scale, w, h, lib, u_right=None):
    """(product Frame, oracle Frame) over the same undistorted keypoints; image bounds as
    Frame::ComputeImageBounds for an undistorted camera (0..cols, 0..rows)."""
    bounds = (0.0, float(w), 0.0, float(h))
    return (orbfe.Frame(kps, desc, scale, bounds, u_right, lib=lib), O.Frame(kps, desc, scale, bounds, u_right))


def check_features_in_area(F, OF, rng, w, h, n=60):
    for _ in range(n):
        x, y = rng.uniform(-20, w + 20), rng.uniform(-20, h + 20)
        r = rng.choice([3.0, 10.0, 40.0, 100.0])
        lo, hi = [(-1, -1), (0, 0), (1, 3), (2, -1), (0, 7)][rng.integers(0, 5)]
        a = F.GetFeaturesInArea(x, y, r, lo, hi)
        b = OF.features_in_area(x, y, r, lo, hi)
        assert np.array_equal(a, b), f"GetFeaturesInArea({x},{y},{r},{lo},{hi})"


def check_search_for_initialization(lib, img1, img2, extract, nfeatures=4000, window=100):
    """config 4: mono frames with nFeatures=4000, SearchForInitialization(F1,F2,prev=F1 pts,100), 0.9."""
    k1, d1 = extract(img1, nfeatures)
    k2, d2 = extract(img2, nfeatures)
    scale = O.Extractor(nfeatures).tables()["scale"]
    h, w = img1.shape
    F1, OF1 = make_frames(k1, d1, scale, w, h, lib)
    F2, OF2 = make_frames(k2, d2, scale, w, h, lib)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = orbfe.OrbMatcher(0.9, True).SearchForInitialization(F1, F2, prev, window)
    on, om12, opm = O.search_for_initialization(OF1, OF2, prev, window, 0.9, True)
    assert n == on, f"SearchForInitialization: {n} vs oracle {on}"
    assert np.array_equal(m12, om12), "vnMatches12 differs"
    assert np.array_equal(pm, opm), "vbPrevMatched differs"
    # second call with the updated vbPrevMatched and no orientation check
    n2, m12b, _ = orbfe.OrbMatcher(0.9, False).SearchForInitialization(F1, F2, pm, window // 2)
    on2, om12b, _ = O.search_for_initialization(OF1, OF2, opm, window // 2, 0.9, False)
    assert n2 == on2 and np.array_equal(m12b, om12b)
    return n


def synth_map_points(kps, desc, rng, n_mp, u_right=None):
    """config 5 recipe: descriptors = random frame descriptors with 0-40 flipped bits,
    projections = that keypoint + jitter, predicted level = octave (+1 w.p. 1/2)."""
    n = len(kps)
    src = rng.integers(0, n, n_mp)
    d = desc[src].copy()
    for i in range(n_mp):
        nb = rng.integers(0, 41)
        bits = rng.choice(256, nb, replace=False)
        for b in bits:
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    lvl = np.minimum(kps["octave"][src] + rng.integers(0, 2, n_mp), 7).astype(np.int32)
    view = rng.choice(np.array([0.999, 0.9], np.float32), n_mp)
    px = (kps["x"][src] + rng.uniform(-3, 3, n_mp)).astype(np.float32)
    py = (kps["y"][src] + rng.uniform(-3, 3, n_mp)).astype(np.float32)
    pxr = (px - rng.uniform(0, 40, n_mp)).astype(np.float32)
    if u_right is not None:
        has = u_right[src] > 0
        pxr[has] = (u_right[src][has] + rng.uniform(-6, 6, has.sum())).astype(np.float32)
    valid = (rng.uniform(0, 1, n_mp) < 0.9).astype(np.uint8)
    has_obs = (rng.uniform(0, 1, n_mp) < 0.8).astype(np.uint8)
    occupied = (rng.uniform(0, 1, n) < 0.1).astype(np.uint8)
    return dict(valid=valid, px=px, py=py, pxr=pxr, lvl=lvl, view=view, desc=d, has_obs=has_obs, occupied=occupied)


def check_search_by_projection_mappoints(lib, kps, desc, scale, w, h, n_mp, seed=0, u_right=None, th=1, nnratio=0.8):
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib, u_right)
    check_features_in_area(F, OF, rng, w, h)
    mp = synth_map_points(kps, desc, rng, n_mp, u_right)
    n, asg = orbfe.OrbMatcher(nnratio).SearchByProjectionMapPoints(F, mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"],
                                                                  mp["view"], mp["desc"], mp["has_obs"], mp["occupied"], th)
    on, oasg = O.search_by_projection_mappoints(OF, mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"],
                                                mp["desc"], mp["has_obs"], mp["occupied"], th, nnratio)
    assert n == on, f"SearchByProjection(map points): {n} vs oracle {on}"
    assert np.array_equal(asg, oasg), "assigned map points differ"
    return n


def check_search_by_projection_lastframe(lib, kps, desc, scale, w, h, seed=0, u_right=None, th=7.0):
    """per-frame tracking match: the 'last frame' points are this frame's keypoints displaced."""
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib, u_right)
    n = len(kps)
    sel = rng.permutation(n)[: max(n // 2, 1)]
    u = (kps["x"][sel] + rng.uniform(-5, 5, len(sel))).astype(np.float32)
    v = (kps["y"][sel] + rng.uniform(-5, 5, len(sel))).astype(np.float32)
    u[:5] -= 2000.0  # out of bounds
    invzc = rng.uniform(0.01, 0.2, len(sel)).astype(np.float32)
    invzc[5:9] = -0.1
    octv = kps["octave"][sel].astype(np.int32)
    ang = ((kps["angle"][sel] + rng.choice([0.0, 0.0, 0.0, 45.0, 170.0], len(sel))) % 360).astype(np.float32)
    d = desc[sel].copy()
    d[:, 0] ^= rng.integers(0, 256, len(sel)).astype(np.uint8)
    valid = (rng.uniform(0, 1, len(sel)) < 0.9).astype(np.uint8)
    has_obs = (rng.uniform(0, 1, len(sel)) < 0.7).astype(np.uint8)
    occupied = (rng.uniform(0, 1, n) < 0.05).astype(np.uint8)
    tot = 0
    for fwd, bwd in ((0, 0), (1, 0), (0, 1)):
        for ori in (True, False):
            m, asg = orbfe.OrbMatcher(0.9, ori).SearchByProjectionLastFrame(F, valid, u, v, invzc, octv, ang, d, has_obs,
                                                                           KITTI["bf"], fwd, bwd, occupied, th)
            om, oasg = O.search_by_projection_lastframe(OF, valid, u, v, invzc, octv, ang, d, has_obs, KITTI["bf"], fwd,
                                                        bwd, occupied, th, ori)
            assert m == om, f"SearchByProjection(last frame) fwd={fwd} bwd={bwd} ori={ori}: {m} vs {om}"
            assert np.array_equal(asg, oasg)
            tot += m
    return tot


def synth_feature_vector(desc, rng, n_nodes=97, noise=0.03):
    """stand-in for DBoW2::FeatureVector (the vocabulary file is a missing blob): node id = a hash of the first
    descriptor bytes, so similar descriptors mostly share a node; a few features are moved to random nodes"""
    node = (desc[:, 0].astype(np.int64) * 7 + (desc[:, 1] >> 5)) % n_nodes
    flip = rng.uniform(0, 1, len(desc)) < noise
    node[flip] = rng.integers(0, n_nodes, flip.sum())
    fv = {}
    for i in rng.permutation(len(desc)) if False else range(len(desc)):
        fv.setdefault(int(node[i]) * 3 + 1, []).append(i)  # sparse, unsorted-looking ids
    return fv


def check_search_by_bow(lib, kps_f, desc_f, kps_kf, desc_kf, scale, w, h, seed=0, nnratio=0.7):
    """SearchByBoW(KeyFrame, Frame): the keyframe side = another frame's features, ~70 % with map points"""
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps_f, desc_f, scale, w, h, lib)
    fv_f = synth_feature_vector(desc_f, rng)
    fv_kf = synth_feature_vector(desc_kf, rng)
    valid = (rng.uniform(0, 1, len(kps_kf)) < 0.7).astype(np.uint8)
    tot = 0
    for ori in (True, False):
        n, m = orbfe.SearchByBoW(F, desc_kf, kps_kf["angle"], valid, fv_kf, fv_f, nnratio, ori)
        on, om = O.search_by_bow(OF, desc_kf, kps_kf["angle"], valid, fv_kf, fv_f, nnratio, ori)
        assert n == on, f"SearchByBoW ori={ori}: {n} vs oracle {on}"
        assert np.array_equal(m, om)
        tot += n
    return tot


# ---- N1: the remaining OrbMatcher searches -----------------------------------------------------------------
def _proj_points(kps, desc, rng, n_mp, jitter=3.0, flip_max=60):
    """map points 'seen' near random keypoints: projection = keypoint + jitter, descriptor = the keypoint's with
    0..flip_max flipped bits, predicted level = octave (+1 w.p. 1/2); ~10 % fail the geometric gates"""
    n = len(kps)
    src = rng.integers(0, n, n_mp)
    d = desc[src].copy()
    for i in range(n_mp):
        for b in rng.choice(256, rng.integers(0, flip_max + 1), replace=False):
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    lvl = np.minimum(kps["octave"][src] + rng.integers(0, 2, n_mp), 7).astype(np.int32)
    u = (kps["x"][src] + rng.uniform(-jitter, jitter, n_mp)).astype(np.float32)
    v = (kps["y"][src] + rng.uniform(-jitter, jitter, n_mp)).astype(np.float32)
    valid = (rng.uniform(0, 1, n_mp) < 0.9).astype(np.uint8)
    return dict(src=src, valid=valid, u=u, v=v, lvl=lvl, desc=d)


def check_search_by_projection_sim3(lib, kps, desc, scale, w, h, n_mp, seed=0, th=10):
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib)
    mp = _proj_points(kps, desc, rng, n_mp)
    matched_in = (rng.uniform(0, 1, len(kps)) < 0.2).astype(np.uint8)
    n, m = orbfe.SearchByProjectionSim3(F, mp["valid"], mp["u"], mp["v"], mp["lvl"], mp["desc"], matched_in, th)
    on, om = O.search_by_projection_sim3(OF, mp["valid"], mp["u"], mp["v"], mp["lvl"], mp["desc"], matched_in, th)
    assert n == on, f"SearchByProjection(KF, Scw): {n} vs oracle {on}"
    assert np.array_equal(m, om)
    return n


def check_search_by_projection_keyframe(lib, kps, desc, scale, w, h, n_kf, seed=0, th=10.0, orb_dist=100):
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib)
    mp = _proj_points(kps, desc, rng, n_kf, jitter=4.0)
    mp["u"][:6] = [-5.0, w + 3.0, 10.0, 10.0, 0.0, float(w)]      # bounds gate (:1490-1495): first two rejected by u
    mp["v"][2:4] = [-1.0, h + 0.5]
    ang = ((kps["angle"][mp["src"]] + rng.choice([0.0, 0.0, 0.0, 30.0, 200.0], n_kf)) % 360).astype(np.float32)
    occupied = (rng.uniform(0, 1, len(kps)) < 0.15).astype(np.uint8)
    tot = 0
    for ori in (True, False):
        for od in (orb_dist, 64):
            n, a = orbfe.SearchByProjectionKeyFrame(F, mp["valid"], mp["u"], mp["v"], mp["lvl"], ang, mp["desc"], occupied, th, od, ori)
            on, oa = O.search_by_projection_keyframe(OF, mp["valid"], mp["u"], mp["v"], mp["lvl"], ang, mp["desc"], occupied, th, od, ori)
            assert n == on, f"SearchByProjection(Frame, KF) ori={ori} ORBdist={od}: {n} vs oracle {on}"
            assert np.array_equal(a, oa)
            tot += n
    return tot


def check_fuse(lib, kps, desc, scale, w, h, n_mp, seed=0, th=3.0, u_right=None):
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib, u_right)
    mp = _proj_points(kps, desc, rng, n_mp, jitter=2.5)
    ur = (mp["u"] - rng.uniform(0, 40, n_mp)).astype(np.float32)
    if u_right is not None:  # most stereo keypoints see a consistent right coordinate, some do not (chi2 gate)
        has = u_right[mp["src"]] >= 0
        ur[has] = (u_right[mp["src"]][has] + rng.uniform(-2.5, 2.5, has.sum())).astype(np.float32)
    tot = 0
    for urs in (ur, None):
        n, b = orbfe.Fuse(F, mp["valid"], mp["u"], mp["v"], urs, mp["lvl"], mp["desc"], th)
        on, ob = O.fuse(OF, mp["valid"], mp["u"], mp["v"], urs, mp["lvl"], mp["desc"], th)
        assert n == on, f"Fuse (ur {'given' if urs is not None else 'none'}): {n} vs oracle {on}"
        assert np.array_equal(b, ob)
        tot += n
    return tot


def check_search_by_sim3(lib, k1, d1, k2, d2, scale, w, h, seed=0, th=7.5, shift=(0.0, 0.0)):
    """KF2 = KF1's scene displaced by `shift`: a map point of KF1 projects at kp - shift in KF2 and back"""
    rng = np.random.default_rng(seed)
    F1, OF1 = make_frames(k1, d1, scale, w, h, lib)
    F2, OF2 = make_frames(k2, d2, scale, w, h, lib)

    def side(k, d, sgn):
        n = len(k)
        dd = d.copy()
        dd[:, 3] ^= rng.integers(0, 4, n).astype(np.uint8)
        u = (k["x"] - sgn * shift[0] + rng.uniform(-2, 2, n)).astype(np.float32)
        v = (k["y"] - sgn * shift[1] + rng.uniform(-2, 2, n)).astype(np.float32)
        lvl = np.minimum(k["octave"] + rng.integers(0, 2, n), 7).astype(np.int32)
        valid = (rng.uniform(0, 1, n) < 0.8).astype(np.uint8)
        return (valid, u, v, lvl, dd)
    s1, s2 = side(k1, d1, 1.0), side(k2, d2, -1.0)
    n, m = orbfe.SearchBySim3(F1, F2, s1, s2, th)
    on, om = O.search_by_sim3(OF1, OF2, s1, s2, th)
    assert n == on, f"SearchBySim3: {n} vs oracle {on}"
    assert np.array_equal(m, om)
    return n


def check_search_by_bow_keyframes(lib, k1, d1, k2, d2, scale, w, h, seed=0, nnratio=0.8):
    rng = np.random.default_rng(seed)
    dup = rng.permutation(len(k2))[: len(k2) // 5]   # exact duplicates => equal distances: the FIRST candidate wins, ratio test fails
    k2, d2 = np.concatenate([k2, k2[dup]]), np.concatenate([d2, d2[dup]])
    F2, OF2 = make_frames(k2, d2, scale, w, h, lib)
    fv1, fv2 = synth_feature_vector(d1, rng), synth_feature_vector(d2, rng)
    v1 = (rng.uniform(0, 1, len(k1)) < 0.7).astype(np.uint8)
    v2 = (rng.uniform(0, 1, len(k2)) < 0.7).astype(np.uint8)
    tot = 0
    for ori in (True, False):
        n, m = orbfe.SearchByBoWKeyFrames(F2, d1, k1["angle"], v1, v2, fv1, fv2, nnratio, ori)
        on, om = O.search_by_bow_keyframes(OF2, d1, k1["angle"], v1, v2, fv1, fv2, nnratio, ori)
        assert n == on, f"SearchByBoW(KF, KF) ori={ori}: {n} vs oracle {on}"
        assert np.array_equal(m, om)
        tot += n
    return tot


def check_search_for_triangulation(lib, k1, d1, k2, d2, scale, w, h, seed=0, shift=(6.0, 0.0)):
    """KF2 = KF1 translated sideways by `shift` pixels: the fundamental matrix of a pure image translation
    (epipolar lines parallel to the shift), plus a generic perturbed one"""
    rng = np.random.default_rng(seed)
    dup = rng.permutation(len(k2))[: len(k2) // 5]   # exact duplicates => equal distances: the LAST candidate must win (:717)
    k2, d2 = np.concatenate([k2, k2[dup]]), np.concatenate([d2, d2[dup]])
    ur2 = np.where(rng.uniform(0, 1, len(k2)) < 0.5, k2["x"] - rng.uniform(1, 50, len(k2)), -1).astype(np.float32)
    F2, OF2 = make_frames(k2, d2, scale, w, h, lib, ur2)
    fv1, fv2 = synth_feature_vector(d1, rng, n_nodes=31), synth_feature_vector(d2, rng, n_nodes=31)
    v1 = (rng.uniform(0, 1, len(k1)) < 0.8).astype(np.uint8)   # 'no map point yet'
    v2 = (rng.uniform(0, 1, len(k2)) < 0.8).astype(np.uint8)
    st1 = (rng.uniform(0, 1, len(k1)) < 0.5).astype(np.uint8)
    tx, ty = shift
    Fa = np.array([[0, 0, -ty], [0, 0, tx], [ty, -tx, 0]], np.float32)          # x2' [t]x x1 = 0 for x2 = x1 + t
    Fb = (Fa + rng.normal(0, 2e-4, (3, 3))).astype(np.float32)
    tot = 0
    for F12, (ex, ey) in ((Fa, (1e6, 188.0)), (Fb, (620.0, 188.0))):
        for only_stereo in (False, True):
            for ori in (True, False):
                n, m = orbfe.SearchForTriangulation(F2, k1, d1, v1, st1, v2, fv1, fv2, F12, ex, ey, only_stereo, ori)
                on, om = O.search_for_triangulation(OF2, k1, d1, v1, st1, v2, fv1, fv2, F12, ex, ey, only_stereo, ori)
                assert n == on, f"SearchForTriangulation stereo={only_stereo} ori={ori}: {n} vs oracle {on}"
                assert np.array_equal(m, om)
                tot += n
    return tot


# ---- N3: vocabulary transform (Frame::ComputeBoW) ------------------------------------------------------------
def synth_vocabulary(rng, k=10, L=3, stop_frac=0.05, prune_frac=0.08, seed_desc=None):
    """a random k-ary tree of depth <= L in the arrays of loadFromTextFile (node 0 = root): a child's descriptor is its
    parent's with ~25 random bits flipped (so descents are decided by small distance gaps and ties do occur), some
    subtrees are pruned early (leaves above level L, as k-means leaves them), some words are stopped (weight 0).  The real
    ORBvoc.txt (k=10, L=6, ~1.08M nodes) is a missing blob; this has the same structure at test size."""
    parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
    frontier = [(0, 0, rng.integers(0, 256, 32, dtype=np.uint8) if seed_desc is None else seed_desc)]
    while frontier:
        nxt = []
        for nid, lvl, d in frontier:
            nchild = k if rng.uniform() > 0.1 else int(rng.integers(2, k + 1))
            for _ in range(nchild):
                cd = d.copy()
                for b in rng.choice(256, int(rng.integers(10, 40)), replace=False):
                    cd[b >> 3] ^= np.uint8(1 << (b & 7))
                cid = len(parent)
                is_leaf = (lvl + 1 == L) or (lvl + 1 >= 1 and rng.uniform() < prune_frac)
                parent.append(nid); leaf.append(1 if is_leaf else 0); desc.append(cd)
                weight.append(0.0 if (is_leaf and rng.uniform() < stop_frac) else (float(rng.uniform(0.1, 9.0)) if is_leaf else 0.0))
                if not is_leaf:
                    nxt.append((cid, lvl + 1, cd))
        frontier = nxt
    # loadFromTextFile appends nodes in file order; parents must precede children (true for ORBvoc.txt and here)
    return (np.array(parent, np.int32), np.array(leaf, np.uint8), np.stack(desc), np.array(weight, np.float64))


def write_vocabulary_text(path, k, L, scoring, weighting, arrays, trailing_newline=False):
    """the ORBvoc.txt format of TemplatedVocabulary::saveToTextFile / loadFromTextFile.  No newline after the last node by
    default: the reference's `while(!f.eof())` loop (TemplatedVocabulary.h:1375) would turn a final empty line into a bogus
    extra child of the root with an uninitialised descriptor (indeterminate); the product loader ignores blank lines."""
    parent, leaf, desc, weight = arrays
    lines = [f"{k} {L} {scoring} {weighting}"]
    for i in range(1, len(parent)):
        lines.append(f"{parent[i]} {leaf[i]} " + " ".join(str(int(b)) for b in desc[i]) + f" {float(weight[i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def assert_bow_equal(got, ref, what=""):
    for key in ("word_id", "node_id"):
        assert np.array_equal(got[key], ref[key]), f"{what}: {key} differs"
    assert np.array_equal(got["bow"][0], ref["bow"][0]), f"{what}: BowVector words differ"
    assert np.array_equal(got["bow"][1], ref["bow"][1]), f"{what}: BowVector values are not bit-identical"
    for j, name in enumerate(("nodes", "start", "idx")):
        assert np.array_equal(got["fv"][j], ref["fv"][j]), f"{what}: FeatureVector {name} differ"


def check_bow_transform(lib, desc, seed=0, k=10, L=3, tmp_path=None):
    """every weighting x the three normalisations, levelsup in (0, 1, 2, L, L+1); optionally through the text file"""
    rng = np.random.default_rng(seed)
    arrays = synth_vocabulary(rng, k, L, seed_desc=desc[0].copy() if len(desc) else None)
    words = 0
    for scoring, weighting in ((0, 0), (1, 0), (5, 0), (0, 1), (5, 1), (0, 2), (1, 3)):
        V = orbfe.OrbVocabulary(k, L, scoring, weighting, *arrays, lib=lib)
        OV = O.Vocabulary(k, L, scoring, weighting, *arrays)
        assert V.info()["nodes"] == len(arrays[0]) and V.info()["words"] == int(arrays[1].sum())
        for levelsup in (0, 1, 2, L, L + 1):
            got, ref = V.transform(desc, levelsup), OV.transform(desc, levelsup)
            assert_bow_equal(got, ref, f"scoring {scoring} weighting {weighting} levelsup {levelsup}")
        words += len(ref["bow"][0])
        if scoring in (0, 1) and len(ref["bow"][1]):
            nrm = np.abs(ref["bow"][1]).sum() if scoring == 0 else np.sqrt((ref["bow"][1] ** 2).sum())
            assert abs(nrm - 1.0) < 1e-9
        V.close()
    if tmp_path is not None:
        path = str(tmp_path / "voc.txt")
        write_vocabulary_text(path, k, L, 0, 0, arrays)
        V = orbfe.OrbVocabulary.loadFromTextFile(path, lib=lib)
        assert_bow_equal(V.transform(desc, 1), O.Vocabulary(k, L, 0, 0, *arrays).transform(desc, 1), "text file")
        V.close()
    return words


def synth_vocabulary_uniform(rng, k=10, L=6, seed_desc=None, flip_bits=24):
    """ORBvoc-sized uniform k-ary tree (k=10, L=6 -> 1 111 111 nodes, 10^6 words) built level by level with numpy;
    same array format as synth_vocabulary"""
    root = rng.integers(0, 256, 32, dtype=np.uint8) if seed_desc is None else seed_desc
    parent, leaf, desc, weight = [np.zeros(1, np.int32)], [np.zeros(1, np.uint8)], [np.zeros((1, 32), np.uint8)], [np.zeros(1)]
    prev_ids, prev_desc, next_id = np.zeros(1, np.int64), root[None, :], 1
    for lvl in range(1, L + 1):
        n = len(prev_ids) * k
        par = np.repeat(prev_ids, k)
        bits = rng.integers(0, 256, (n, flip_bits))
        mask = np.zeros((n, 32), np.uint8)
        np.bitwise_or.at(mask, (np.repeat(np.arange(n), flip_bits), (bits >> 3).ravel()), (1 << (bits & 7)).astype(np.uint8).ravel())
        d = np.repeat(prev_desc, k, axis=0) ^ mask
        ids = np.arange(next_id, next_id + n, dtype=np.int64)
        # node ids must follow loadFromTextFile order (a parent precedes its children): level order satisfies it
        parent.append(par.astype(np.int32)); desc.append(d)
        is_leaf = lvl == L
        leaf.append(np.full(n, 1 if is_leaf else 0, np.uint8))
        weight.append(rng.uniform(0.1, 12.0, n) * (rng.uniform(0, 1, n) > 0.01) if is_leaf else np.zeros(n))
        prev_ids, prev_desc, next_id = ids, d, next_id + n
    return (np.concatenate(parent), np.concatenate(leaf), np.concatenate(desc), np.concatenate(weight).astype(np.float64))


# ---- N2: the Frame tail ------------------------------------------------------------------------------------------
def check_undistort_keypoints(lib, kps, seed=0):
    rng = np.random.default_rng(seed)
    cams = [(718.856, 718.856, 607.1928, 185.2157, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
            (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]),
            (458.654, 457.296, 367.215, 248.375, [-0.2834, 0.0739, 0.00019, 1.76e-05, 0.0, 0.01, -0.02, 0.003]),
            (400.0, 410.0, 320.0, 240.0, [0.9, -2.0, 0.01, 0.01, 5.0]),      # reaches the icdist < 0 branch
            (718.856, 718.856, 607.1928, 185.2157, [0.0, 0.0, 0.0, 0.0])]     # KITTI: identity short-circuit
    k = kps.copy()
    k["x"] += rng.uniform(-0.5, 0.5, len(k)).astype(np.float32)  # sub-pixel coordinates as well
    for fx, fy, cx, cy, dist in cams:
        got = orbfe.UndistortKeyPoints(k, fx, fy, cx, cy, dist, lib=lib)
        ref = O.undistort_points(np.stack([k["x"], k["y"]], 1), fx, fy, cx, cy, dist)
        assert np.array_equal(got["x"], ref[:, 0]) and np.array_equal(got["y"], ref[:, 1]), f"undistort {dist}"
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(got[f], k[f])
    return len(k)


def synth_local_map(rng, n, fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, w=1241, h=376):
    """a camera pose + n local map points: most in front of the camera and inside the image, some behind it, outside
    the image, out of the scale-invariance range, or seen from too oblique an angle"""
    ang = rng.uniform(-0.2, 0.2, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(np.float32)
    t = rng.normal(0, 2, 3).astype(np.float32)
    Ow = (-R.T.astype(np.float64) @ t.astype(np.float64)).astype(np.float32)
    z = rng.uniform(-5, 80, n)
    x = (rng.uniform(-0.15 * w, 1.15 * w, n) - cx) / fx * z
    y = (rng.uniform(-0.15 * h, 1.15 * h, n) - cy) / fy * z
    Pc = np.stack([x, y, z], 1)
    Pw = ((Pc - t) @ R.astype(np.float64)).astype(np.float32)      # Pw = R^T (Pc - t)
    PO = Pw - Ow
    d = np.linalg.norm(PO, axis=1)
    nrm = PO / np.maximum(d[:, None], 1e-6) + rng.normal(0, 0.6, (n, 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    maxd = (d * rng.uniform(0.7, 4.0, n)).astype(np.float32)
    mind = (maxd / np.float32(1.2 ** 7) * rng.uniform(0.5, 1.2, n)).astype(np.float32)
    # some points sit EXACTLY on a level boundary of PredictScale: ratio = 1.2^k
    kk = rng.integers(0, 8, n)
    onb = rng.uniform(0, 1, n) < 0.05
    maxd[onb] = (d[onb].astype(np.float32) * np.float32(1.2) ** kk[onb].astype(np.float32)).astype(np.float32)
    # the reference holds max_dist_ / min_dist_ and exposes 1.2f * max_dist_ / 0.8f * min_dist_ (map_point.cpp:356-364)
    raw = maxd
    maxd, mind = (np.float32(1.2) * raw).astype(np.float32), (np.float32(0.8) * mind).astype(np.float32)
    return dict(world=Pw, normal=nrm, min_dist=mind, max_dist=maxd, max_dist_raw=raw, Rcw=R, tcw=t, Ow=Ow, fx=fx, fy=fy, cx=cx, cy=cy,
                bounds=(0.0, float(w), 0.0, float(h)))


def check_is_in_frustum(lib, n=20000, seed=0):
    rng = np.random.default_rng(seed)
    m = synth_local_map(rng, n)
    lsf = float(np.log(np.float32(1.2)).astype(np.float32))   # log_scale_factor_ = log(scale_factor_) as float
    args = (m["world"], m["normal"], m["min_dist"], m["max_dist"], m["max_dist_raw"], m["Rcw"], m["tcw"], m["Ow"], m["fx"], m["fy"], m["cx"], m["cy"],
            KITTI["bf"], m["bounds"], lsf, 8, 0.5)
    cnt, got = orbfe.IsInFrustum(*args, lib=lib)
    ocnt, ref = O.is_in_frustum(*args)
    assert cnt == ocnt, f"IsInFrustum: {cnt} vs oracle {ocnt}"
    for key in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos"):
        assert np.array_equal(got[key], ref[key]), f"IsInFrustum: {key} differs"
    assert 0.1 * n < cnt < 0.9 * n
    assert len(np.unique(ref["level"][ref["in_view"] == 1])) >= 6
    return cnt


def check_logf(lib, n=400000, seed=0):
    """device restatement of glibc logf == this machine's libm logf (what PredictScale's std::log(float) calls)"""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    rng = np.random.default_rng(seed)
    bits = rng.integers(0x00800000, 0x7f800000, n, dtype=np.uint32)           # every positive normal exponent
    x = np.concatenate([bits.view(np.float32), rng.uniform(0.05, 40.0, n).astype(np.float32),
                        (np.float32(1.2) ** np.arange(-8, 9)).astype(np.float32), np.array([1.0, 1e-40, 3e-39], np.float32)])
    ref = np.zeros_like(x)
    # vectorised libm call: logf through numpy would use its own SIMD log, so call libm per element in C via ctypes arrays
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    step = max(1, len(x) // 60000)   # a 60k-element sample keeps the ctypes loop short; the full set goes through the oracle
    idx = np.arange(0, len(x), step)
    ref_s = np.array([libm.logf(float(v)) for v in x[idx]], np.float32)
    got = orbfe.debug_logf(x, lib=lib)
    assert np.array_equal(got[idx].view(np.uint32), ref_s.view(np.uint32)), "logf restatement differs from libm"
    return len(idx)


def check_search_local_points(lib, kps, desc, scale, w, h, seed=0, u_right=None, th=1, n_extra=2000):
    """Tracker::SearchLocalPoints in one device-resident call == oracle IsInFrustum followed by oracle SearchByProjection"""
    rng = np.random.default_rng(seed)
    F, OF = make_frames(kps, desc, scale, w, h, lib, u_right)
    fx = fy = np.float32(718.856); cx, cy = np.float32(607.1928), np.float32(185.2157)
    n0 = len(kps)
    src = np.concatenate([np.arange(n0), rng.integers(0, n0, n_extra)])
    n = len(src)
    z = rng.uniform(3, 70, n).astype(np.float32)
    z[rng.uniform(0, 1, n) < 0.03] = -4.0
    jx, jy = rng.uniform(-2.5, 2.5, n).astype(np.float32), rng.uniform(-2.5, 2.5, n).astype(np.float32)
    world = np.stack([(kps["x"][src] + jx - cx) / fx * z, (kps["y"][src] + jy - cy) / fy * z, z], 1).astype(np.float32)
    t = np.array([0.03, -0.01, 0.2], np.float32)
    R = np.eye(3, dtype=np.float32)
    Ow = (-t).astype(np.float32)
    PO = world - Ow
    d = np.linalg.norm(PO, axis=1).astype(np.float32)
    nrm = (PO / np.maximum(d[:, None], 1e-6) + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    raw = (d * scale[np.minimum(kps["octave"][src] + rng.integers(0, 2, n), len(scale) - 1)]).astype(np.float32)
    maxd = (np.float32(1.2) * raw).astype(np.float32)
    mind = (np.float32(0.8) * raw / scale[-1]).astype(np.float32)
    md = desc[src].copy()
    for i in range(n):
        for b in rng.choice(256, rng.integers(0, 40), replace=False):
            md[i, b >> 3] ^= np.uint8(1 << (b & 7))
    has_obs = (rng.uniform(0, 1, n) < 0.8).astype(np.uint8)
    occupied = (rng.uniform(0, 1, n0) < 0.1).astype(np.uint8)
    lsf = float(np.log(np.float32(1.2)).astype(np.float32))
    bf = KITTI["bf"]
    nv, nm, in_view, lvl, asg = orbfe.SearchLocalPoints(F, world, nrm, mind, maxd, raw, R, t, Ow, float(fx), float(fy), float(cx), float(cy),
                                                        bf, lsf, md, has_obs, occupied, th, 0.8)
    onv, tr = O.is_in_frustum(world, nrm, mind, maxd, raw, R, t, Ow, float(fx), float(fy), float(cx), float(cy), bf,
                              (0.0, float(w), 0.0, float(h)), lsf, len(scale), 0.5)
    onm, oasg = O.search_by_projection_mappoints(OF, tr["in_view"], tr["proj_x"], tr["proj_y"], tr["proj_xr"], tr["level"], tr["view_cos"],
                                                 md, has_obs, occupied, th, 0.8)
    assert nv == onv and np.array_equal(in_view, tr["in_view"]), f"in view: {nv} vs oracle {onv}"
    assert np.array_equal(lvl[in_view == 1], tr["level"][in_view == 1])
    assert nm == onm, f"SearchLocalPoints: {nm} matches vs oracle {onm}"
    assert np.array_equal(asg, oasg)
    return nm


def check_empty_inputs(lib, ka, da, scale):
    """zero map points / queries / features through every N1-N3 entry point: no crash, empty results"""
    F = orbfe.Frame(ka, da, scale, (0, 640, 0, 200), lib=lib)
    E = orbfe.Frame(ka[:0], da[:0], scale, (0, 640, 0, 200), lib=lib)
    z8, zf, zi = np.zeros(0, np.uint8), np.zeros(0, np.float32), np.zeros(0, np.int32)
    zd = np.zeros((0, 32), np.uint8)
    n, m = orbfe.SearchByProjectionSim3(F, z8, zf, zf, zi, zd, np.zeros(len(ka), np.uint8), 10)
    assert n == 0 and (m == -1).all()
    n, m = orbfe.SearchByProjectionKeyFrame(F, z8, zf, zf, zi, zf, zd, np.zeros(len(ka), np.uint8), 10.0, 100)
    assert n == 0 and (m == -1).all()
    for ur in (zf, None):
        n, b = orbfe.Fuse(F, z8, zf, zf, ur, zi, zd, 3.0)
        assert n == 0 and len(b) == 0
    side0 = (z8, zf, zf, zi, zd)
    sideF = (np.zeros(len(ka), np.uint8), np.zeros(len(ka), np.float32), np.zeros(len(ka), np.float32), np.zeros(len(ka), np.int32), da)
    assert orbfe.SearchBySim3(E, F, side0, sideF, 7.5)[0] == 0
    n, m = orbfe.SearchBySim3(F, E, sideF, side0, 7.5)
    assert n == 0 and (m == -1).all()
    assert orbfe.SearchByBoW(F, zd, zf, z8, {}, {}, 0.7, True)[0] == 0
    assert orbfe.SearchByBoW(E, da, ka["angle"], np.ones(len(ka), np.uint8), {3: [0, 1]}, {}, 0.7, True)[0] == 0
    n, m = orbfe.SearchByBoWKeyFrames(F, zd, zf, z8, np.ones(len(ka), np.uint8), {}, {1: [0]}, 0.8, True)
    assert n == 0 and len(m) == 0
    n, m = orbfe.SearchForTriangulation(F, ka[:0], zd, z8, z8, np.ones(len(ka), np.uint8), {}, {}, np.eye(3), 0.0, 0.0)
    assert n == 0 and len(m) == 0
    cnt, out = orbfe.IsInFrustum(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), zf, zf, zf, np.eye(3), np.zeros(3), np.zeros(3),
                                 700.0, 700.0, 600.0, 180.0, 380.0, (0, 640, 0, 200), 0.18, 8, lib=lib)
    assert cnt == 0 and len(out["in_view"]) == 0
    nv, nm, iv, lv, asg = orbfe.SearchLocalPoints(F, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), zf, zf, zf, np.eye(3), np.zeros(3),
                                                  np.zeros(3), 700.0, 700.0, 600.0, 180.0, 380.0, 0.18, zd, z8, np.zeros(len(ka), np.uint8))
    assert nv == 0 and nm == 0 and (asg == -1).all()
    assert len(orbfe.debug_logf(zf, lib=lib)) == 0


def check_frame_from_extractor(lib, left, right, nfeatures=2000, seed=0):
    """orbfe_frame_from_extractor: the matcher view built device to device from an extractor slot behaves exactly like the one
    built from the downloaded host arrays (and like the oracle's)"""
    rng = np.random.default_rng(seed)
    h, w = left.shape
    ex = orbfe.ORBextractor(nfeatures, lib=lib, max_images=2)
    ex.upload([left, right])
    ex.run(2)
    ex.run_stereo(1, KITTI["bf"], KITTI["bf"] / KITTI["fx"])
    b = ex.download(2, ex.make_buffers(2, stereo=True))
    n = int(b["n"][0])
    kps, desc, ur = b["kps"][0, :n].copy(), b["desc"][0, :n].copy(), b["ur"][0, :n].copy()
    scale = ex.GetScaleFactors()
    bounds = (0.0, float(w), 0.0, float(h))
    FD = orbfe.Frame.from_extractor(ex, bounds, slot=0, stereo=True)
    FR = orbfe.Frame.from_extractor(ex, bounds, slot=1, stereo=False)
    assert len(FD.kps) == n and len(FR.kps) == int(b["n"][1])
    FH, OF = make_frames(kps, desc, scale, w, h, lib, ur)
    check_features_in_area(FD, OF, rng, w, h)
    mp = synth_map_points(kps, desc, rng, 4000, ur)
    args = (mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"], mp["desc"], mp["has_obs"], mp["occupied"])
    m = orbfe.OrbMatcher(0.8)
    nd, ad = m.SearchByProjectionMapPoints(FD, *args, 1)
    nh, ah = m.SearchByProjectionMapPoints(FH, *args, 1)
    on, oa = O.search_by_projection_mappoints(OF, *args, 1, 0.8)
    assert nd == nh == on and np.array_equal(ad, ah) and np.array_equal(ad, oa)
    # right-image slot, monocular view: SearchForInitialization between the two device-built frames == host-built ones
    nr = int(b["n"][1])
    kr, dr = b["kps"][1, :nr].copy(), b["desc"][1, :nr].copy()
    FD0 = orbfe.Frame.from_extractor(ex, bounds, slot=0, stereo=False)
    FH0, OF0 = make_frames(kps, desc, scale, w, h, lib)
    FHR, OFR = make_frames(kr, dr, scale, w, h, lib)
    prev = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
    r1 = orbfe.OrbMatcher(0.9, True).SearchForInitialization(FD0, FR, prev, 100)
    r2 = orbfe.OrbMatcher(0.9, True).SearchForInitialization(FH0, FHR, prev, 100)
    r3 = O.search_for_initialization(OF0, OFR, prev, 100, 0.9, True)
    assert r1[0] == r2[0] == r3[0] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[1], r3[1]) and np.array_equal(r1[2], r3[2])
    # per-frame form: one handle refreshed in place -- first a nearly empty frame (few keypoints: small arrays), then the rich
    # one (the arrays and the lazily sized resolve buffers must grow), then the right image (same capacity, no allocation)
    flat = (left.astype(np.float32) * 0.02 + 120).astype(np.uint8)
    ex.upload([flat, flat]); ex.run(2)
    FT = orbfe.Frame.from_extractor(ex, bounds, slot=0)
    n_flat = len(FT.kps)
    m.SearchByProjectionMapPoints(FT, *[a[:10] if i < 8 else np.zeros(n_flat, np.uint8) for i, a in enumerate(args)], 1)  # sizes its resolve buffers
    for imgs, slot, (kk, dd) in (([left, right], 0, (kps, desc)), ([left, right], 1, (kr, dr))):
        ex.upload(imgs); ex.run(2)
        FT.refresh_from_extractor(ex, slot=slot)
        assert len(FT.kps) == len(kk) and len(kk) > n_flat
        FHt, OFt = make_frames(kk, dd, scale, w, h, lib)
        check_features_in_area(FT, OFt, rng, w, h, n=30)
        mpt = synth_map_points(kk, dd, rng, 3000)
        at = (mpt["valid"], mpt["px"], mpt["py"], mpt["pxr"], mpt["lvl"], mpt["view"], mpt["desc"], mpt["has_obs"], mpt["occupied"])
        n1_, a1_ = m.SearchByProjectionMapPoints(FT, *at, 1)
        on_, oa_ = O.search_by_projection_mappoints(OFt, *at, 1, 0.8)
        assert n1_ == on_ and np.array_equal(a1_, oa_)
    ex.close()
    return nd
