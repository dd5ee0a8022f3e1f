"""Parity tests proper: the sm_100a build of liborbfe.so, called through the C ABI, against the CPU
oracle on the same seeded inputs (BASELINE.json configs 1-3), plus size-independent properties at
full batch size.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import parity_common as P
from slam_framework_b200 import orbfe, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = orbfe.load()
    if L.orbfe_device_count() < 1:
        pytest.fail("no CUDA device: the product library has no CPU path")
    assert b"EMULATED" not in L.orbfe_version()
    return L


def test_config1_single_kitti_frame(lib):
    kps, desc = P.check_extract(lib, synth.frame(seed=0))
    assert 1900 < len(kps) <= 2000 + 24


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_extract_more_frames(lib, seed):
    P.check_extract(lib, synth.frame(seed=seed))


@pytest.mark.parametrize("h,w,nf", [(120, 400, 500), (97, 131, 300), (480, 640, 1000), (1080, 1920, 8000)])
def test_extract_other_sizes(lib, h, w, nf):
    P.check_extract(lib, synth.frame(h, w, seed=h), nfeatures=nf)


def test_extract_dense_noise_and_strided(lib):
    rng = np.random.default_rng(4)
    big = rng.integers(0, 256, (400, 1300), dtype=np.uint8)
    P.check_extract(lib, big[10:386, 20:1261])  # white noise: ~65k candidates, non-contiguous rows


@pytest.mark.parametrize("pct", [1, 5, 30])
def test_fast_dense_form_when_the_candidate_queue_overflows(lib, monkeypatch, pct):
    """Same as the emulation test: the queue-overflow (dense) form of k_fast_cells must give identical keypoints."""
    monkeypatch.setenv("ORBFE_TEST_FAST_QUEUE_PCT", str(pct))
    P.check_extract(lib, synth.frame(seed=5))
    rng = np.random.default_rng(9)
    P.check_extract(lib, rng.integers(0, 256, (376, 1241), dtype=np.uint8))
    low = (synth.frame(seed=3) // 8 + 100).astype(np.uint8)
    P.check_extract(lib, low)
    left, right = synth.stereo_pair(seed=6)
    P.check_stereo(lib, left, right)


def test_flat_and_empty_images(lib):
    ex = orbfe.ORBextractor(lib=lib)
    kps, desc = ex.Compute(np.full((376, 1241), 90, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    kps, desc = ex.Compute(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0


def test_handle_reuse_across_sizes(lib):
    ex = orbfe.ORBextractor(lib=lib)
    import oracle_lib as O
    for shape in ((376, 1241), (200, 640), (376, 1241)):
        img = synth.frame(*shape, seed=11)
        kps, desc = ex.Compute(img)
        ok, od = O.Extractor().extract(img)
        P.assert_kps_equal(kps, ok)
        assert np.array_equal(desc, od)


@pytest.mark.parametrize("seed", [0, 5])
def test_config2_stereo_pair(lib, seed):
    left, right = synth.stereo_pair(seed=seed)
    assert P.check_stereo(lib, left, right) > 300


def test_config3_batch_sharded_stereo(lib):
    pairs = [synth.stereo_pair(seed=100 + s) for s in range(6)]
    assert P.check_batch_stereo(lib, pairs) > 6 * 300


def test_batch_is_deterministic_and_slot_independent(lib):
    """size-independent property at a larger batch: the same pair in every slot position gives
    byte-identical outputs, run after run."""
    l, r = synth.stereo_pair(seed=42)
    n = 32
    ex = orbfe.ORBextractor(lib=lib, max_images=2 * n)
    ex.upload([l, r] * n)
    outs = []
    for _ in range(2):
        ex.run(2 * n)
        ex.run_stereo(n, P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"])
        b = ex.download(2 * n, ex.make_buffers(2 * n, stereo=True))
        outs.append(b)
    # slot 0 against the oracle: this batch size runs the 32-keypoints-per-warp descriptor path
    import oracle_lib as O
    oL, oR = O.Extractor(), O.Extractor()
    okl, odl = oL.extract(l)
    okr, odr = oR.extract(r)
    P.assert_kps_equal(outs[0]["kps"][0, :outs[0]["n"][0]], okl, "slot 0")
    assert np.array_equal(outs[0]["desc"][0, :outs[0]["n"][0]], odl)
    _, our, _ = O.stereo_match(oL, oR, okl, odl, okr, odr, P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"])
    assert np.array_equal(outs[0]["ur"][0, :outs[0]["n"][0]], our)
    for b in outs:
        n0 = b["n"][0]
        for p in range(n):
            assert b["n"][2 * p] == n0
            assert np.array_equal(b["kps"][2 * p, :n0], outs[0]["kps"][0, :n0])
            assert np.array_equal(b["desc"][2 * p, :n0], outs[0]["desc"][0, :n0])
            assert np.array_equal(b["ur"][2 * p, :n0], outs[0]["ur"][0, :n0])


def test_two_handles_concurrent_threads(lib):
    """frame.cpp:86-89 runs the left and right extractor on two std::threads."""
    import threading
    import oracle_lib as O
    l, r = synth.stereo_pair(seed=9)
    exs = [orbfe.ORBextractor(lib=lib), orbfe.ORBextractor(lib=lib)]
    res = [None, None]

    def work(i, img):
        for _ in range(5):
            res[i] = exs[i].Compute(img)
    ts = [threading.Thread(target=work, args=(i, im)) for i, im in enumerate((l, r))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for (k, d), img in zip(res, (l, r)):
        ok, od = O.Extractor().extract(img)
        P.assert_kps_equal(k, ok)
        assert np.array_equal(d, od)


# ---- matchers: configs 4 and 5 + per-frame tracking match -------------------------------------------
def gpu_extract(lib):
    def f(img, nfeatures):
        ex = orbfe.ORBextractor(nfeatures, lib=lib)
        out = ex.Compute(img)
        ex.close()
        return out
    return f


def test_descriptor_distance(lib):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (20000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (20000, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10:20] = ~a[10:20]
    d = orbfe.DescriptorDistance(a, b, lib=lib)
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(1))


def test_config4_search_for_initialization(lib):
    a, b = synth.shifted_frame(21, dx=8, dy=4)
    assert P.check_search_for_initialization(lib, a, b, gpu_extract(lib), nfeatures=4000) > 100
    a, b = synth.shifted_frame(22, dx=-8, dy=-4)
    assert P.check_search_for_initialization(lib, a, b, gpu_extract(lib), nfeatures=4000) > 100


@pytest.mark.parametrize("h,w", [(1080, 1920), (2160, 3840)])
def test_config5_search_by_projection_20k_mappoints(lib, h, w):
    import oracle_lib as O
    img = synth.frame(h, w, seed=w)
    kps, desc = P.check_extract(lib, img, nfeatures=8000, stages=False)
    scale = O.Extractor(8000).tables()["scale"]
    assert P.check_search_by_projection_mappoints(lib, kps, desc, scale, w, h, 20000, seed=3) > 3000


def test_tracking_search_by_projection_lastframe(lib):
    import oracle_lib as O
    l, r = synth.stereo_pair(seed=31)
    eL, eR = orbfe.ORBextractor(lib=lib), orbfe.ORBextractor(lib=lib)
    kl, dl = eL.Compute(l)
    kr, dr = eR.Compute(r)
    _, ur, _ = orbfe.ComputeStereoMatches(eL, eR, kl, dl, kr, dr, P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"])
    scale = eL.GetScaleFactors()
    assert P.check_search_by_projection_lastframe(lib, kl, dl, scale, 1241, 376, seed=8, u_right=ur, th=7.0) > 1000
    assert P.check_search_by_projection_mappoints(lib, kl, dl, scale, 1241, 376, 5000, seed=9, u_right=ur) > 500


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_colour_frames_converted_on_device(lib, channels, rgb):
    """N4 (tracker.cpp:110-127): KITTI image_2/image_3 are colour PNGs; gray conversion on the GPU"""
    import oracle_lib as O
    rng = np.random.default_rng(channels)
    gray = synth.frame(seed=12).astype(np.int32)
    col = np.clip(gray[..., None] + rng.integers(-25, 26, (376, 1241, channels)), 0, 255).astype(np.uint8)
    ex = orbfe.ORBextractor(lib=lib, max_images=2)
    ex.upload_color([col, col[:, ::-1].copy()], rgb=rgb)
    ex.run(2)
    b = ex.download(2, ex.make_buffers(2))
    ref = O.cvt_gray(col, rgb)
    try:
        import cv2
        code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY,
                (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)]
        assert np.array_equal(cv2.cvtColor(col, code), ref)
    except ImportError:
        pass
    assert np.array_equal(ex.pyramid_level(0, 0), ref)
    ok, od = O.Extractor().extract(ref)
    P.assert_kps_equal(b["kps"][0, :b["n"][0]], ok)
    assert np.array_equal(b["desc"][0, :b["n"][0]], od)


def test_n1_search_by_bow_keyframe_frame(lib):
    """SearchByBoW(KeyFrame, Frame) on the same kernels (TrackReferenceKeyFrame, tracker.cpp:657-694)"""
    a, b = synth.shifted_frame(41, dx=6, dy=3)
    ex = gpu_extract(lib)
    ka, da = ex(a, 2000)
    kb, db = ex(b, 2000)
    import oracle_lib as O
    scale = O.Extractor().tables()["scale"]
    assert P.check_search_by_bow(lib, kb, db, ka, da, scale, 1241, 376, seed=3) > 200
    assert P.check_search_by_bow(lib, ka, da, ka, da, scale, 1241, 376, seed=4, nnratio=0.75) > 1000


@pytest.fixture(scope="module")
def kitti_two_frames(lib):
    import oracle_lib as O
    a, b = synth.shifted_frame(43, dx=6, dy=0)
    ex = gpu_extract(lib)
    ka, da = ex(a, 2000)
    kb, db = ex(b, 2000)
    return ka, da, kb, db, O.Extractor().tables()["scale"]


def test_n1_search_by_projection_loop_and_relocalisation(lib, kitti_two_frames):
    """SearchByProjection(KF, Scw, ...) (loop closing, :384-497) and SearchByProjection(Frame, KF, ...) (relocalisation,
    :1455-1582)"""
    ka, da, _, _, scale = kitti_two_frames
    assert P.check_search_by_projection_sim3(lib, ka, da, scale, 1241, 376, 6000, seed=51) > 400
    assert P.check_search_by_projection_keyframe(lib, ka, da, scale, 1241, 376, 2500, seed=52) > 1000


def test_n1_fuse_both_overloads(lib, kitti_two_frames):
    """the searches of Fuse (:804-954 with the stereo chi2 gate, :956-1079 without)"""
    ka, da, _, _, scale = kitti_two_frames
    rng = np.random.default_rng(53)
    ur = np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - rng.uniform(1, 60, len(ka)), -1).astype(np.float32)
    assert P.check_fuse(lib, ka, da, scale, 1241, 376, 8000, seed=54, u_right=ur) > 1000


def test_n1_search_by_sim3(lib, kitti_two_frames):
    ka, da, kb, db, scale = kitti_two_frames
    assert P.check_search_by_sim3(lib, ka, da, kb, db, scale, 1241, 376, seed=55, shift=(6.0, 0.0)) > 300


def test_n1_search_by_bow_keyframes_and_triangulation(lib, kitti_two_frames):
    ka, da, kb, db, scale = kitti_two_frames
    assert P.check_search_by_bow_keyframes(lib, ka, da, kb, db, scale, 1241, 376, seed=56) > 150
    assert P.check_search_for_triangulation(lib, ka, da, kb, db, scale, 1241, 376, seed=57) > 300


def test_n3_bow_transform_orbvoc_sized_vocabulary(lib, kitti_two_frames, tmp_path):
    """Frame::ComputeBoW = DBoW2 transform(features, BowVector, FeatureVector, 4) on a k=10, L=6 tree
    (the shape of ORBvoc.txt, which is a missing blob) + the small irregular trees and the text-file loader"""
    import oracle_lib as O
    ka, da, kb, db, _ = kitti_two_frames
    rng = np.random.default_rng(61)
    arrays = P.synth_vocabulary_uniform(rng, 10, 6, seed_desc=da[0].copy())
    assert len(arrays[0]) == 1111111
    V = orbfe.OrbVocabulary(10, 6, 0, 0, *arrays, lib=lib)
    OV = O.Vocabulary(10, 6, 0, 0, *arrays)
    for d in (da, db):
        got, ref = V.transform(d, 4), OV.transform(d, 4)
        P.assert_bow_equal(got, ref, "ORBvoc-sized")
        assert len(ref["bow"][0]) > 100 and len(ref["fv"][0]) > 20
    # the FeatureVectors drive SearchByBoW exactly as Tracking::TrackReferenceKeyFrame does (tracker.cpp:657-694)
    fa, fb = orbfe.feature_vector_dict(V.transform(da, 4)["fv"]), orbfe.feature_vector_dict(V.transform(db, 4)["fv"])
    scale = O.Extractor().tables()["scale"]
    F, OF = P.make_frames(kb, db, scale, 1241, 376, lib)
    valid = np.ones(len(ka), np.uint8)
    n, m = orbfe.SearchByBoW(F, da, ka["angle"], valid, fa, fb, 0.7, True)
    on, om = O.search_by_bow(OF, da, ka["angle"], valid, fa, fb, 0.7, True)
    assert n == on and np.array_equal(m, om)
    V.close()
    assert P.check_bow_transform(lib, da, seed=62, k=10, L=3, tmp_path=tmp_path) > 1000
    assert P.check_bow_transform(lib, db[:500], seed=63, k=4, L=5) > 300


def test_n2_frame_tail(lib, kitti_two_frames):
    """UndistortKeyPoints (cv::undistortPoints, pinned to cv2), IsInFrustum over 50k local map points with
    PredictScale's logf, and the frustum -> SearchByProjection pipeline of Tracker::SearchLocalPoints"""
    import oracle_lib as O
    ka, da, _, _, scale = kitti_two_frames
    assert P.check_undistort_keypoints(lib, ka, seed=71) > 1500
    assert P.check_is_in_frustum(lib, 50000, seed=72) > 5000
    assert P.check_logf(lib, 400000, seed=73) > 50000
    # SearchLocalPoints: map points = the frame's own keypoints back-projected at random depths, seen from a nearby pose
    rng = np.random.default_rng(74)
    n = len(ka)
    fx = fy = 718.856; cx, cy = 607.1928, 185.2157
    z = rng.uniform(4, 60, n)
    Pw = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    R = np.eye(3, dtype=np.float32)
    t = np.array([0.05, -0.02, 0.1], np.float32)
    Ow = (-t).astype(np.float32)
    nrm = (-(Pw - Ow) / np.linalg.norm(Pw - Ow, axis=1, keepdims=True)).astype(np.float32) * np.float32(-1.0)
    d = np.linalg.norm(Pw - Ow, axis=1).astype(np.float32)
    maxd = (d * np.float32(1.2) ** ka["octave"].astype(np.float32)).astype(np.float32) * np.float32(1.05)
    mind = (maxd / np.float32(1.2 ** 8)).astype(np.float32)
    lsf = float(np.log(np.float32(1.2)).astype(np.float32))
    args = (Pw, nrm, mind, (np.float32(1.2) * maxd).astype(np.float32), maxd, R, t, Ow, fx, fy, cx, cy, P.KITTI["bf"], (0.0, 1241.0, 0.0, 376.0), lsf, 8, 0.5)
    cnt, tr = orbfe.IsInFrustum(*args, lib=lib)
    ocnt, otr = O.is_in_frustum(*args)
    assert cnt == ocnt and cnt > 0.8 * n
    for key in tr:
        assert np.array_equal(tr[key], otr[key]), key
    F, OF = P.make_frames(ka, da, scale, 1241, 376, lib)
    has_obs, occ = np.ones(n, np.uint8), np.zeros(n, np.uint8)
    m = orbfe.OrbMatcher(0.8)
    nm, asg = m.SearchByProjectionMapPoints(F, tr["in_view"], tr["proj_x"], tr["proj_y"], tr["proj_xr"], tr["level"], tr["view_cos"],
                                            da, has_obs, occ, 1)
    onm, oasg = O.search_by_projection_mappoints(OF, otr["in_view"], otr["proj_x"], otr["proj_y"], otr["proj_xr"], otr["level"],
                                                 otr["view_cos"], da, has_obs, occ, 1, 0.8)
    assert nm == onm and np.array_equal(asg, oasg) and nm > 0.5 * n


def test_config3_full_length_sequence_sharded_properties(lib):
    """BASELINE config 3 at its full size: 4541 stereo pairs (KITTI-00 length) cut into 8 rank shards and processed in
    64-pair batches through the device-resident path.  Size-independent properties: every pair's outputs depend only on
    its content (same digest wherever it lands in a batch or shard), the shards cover every index exactly once, and a
    sample of pairs equals the oracle."""
    import oracle_lib as O
    from slam_framework_b200 import shard
    n_total, distinct, B = 4541, 40, 64
    pairs = [synth.stereo_pair(seed=300 + s) for s in range(distinct)]
    ex = orbfe.ORBextractor(lib=lib, max_images=2 * B)
    bf, base = P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"]
    seen, first, total_matches = set(), {}, 0
    for rank in range(8):
        lo, hi = shard.shard_range(n_total, rank, 8)
        for b0 in range(lo, hi, B):
            idx = list(range(b0, min(b0 + B, hi)))
            ex.upload([im for i in idx for im in pairs[i % distinct]])
            ex.run(2 * len(idx))
            ex.run_stereo(len(idx), bf, base)
            out = ex.download(2 * len(idx), ex.make_buffers(2 * len(idx), stereo=True))
            for p, i in enumerate(idx):
                assert i not in seen
                seen.add(i)
                nl, nr = out["n"][2 * p], out["n"][2 * p + 1]
                dg = shard.digest([out["kps"][2 * p, :nl], out["desc"][2 * p, :nl], out["kps"][2 * p + 1, :nr],
                                   out["desc"][2 * p + 1, :nr], out["ur"][2 * p, :nl], out["depth"][2 * p, :nl]])
                key = i % distinct
                if key not in first:
                    first[key] = dg
                    if key % 13 == 0:  # oracle spot checks
                        oL, oR = O.Extractor(), O.Extractor()
                        okl, odl = oL.extract(pairs[key][0]); okr, odr = oR.extract(pairs[key][1])
                        P.assert_kps_equal(out["kps"][2 * p, :nl], okl, f"pair {i} left")
                        assert np.array_equal(out["desc"][2 * p + 1, :nr], odr)
                        _, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, base)
                        assert np.array_equal(out["ur"][2 * p, :nl], our) and np.array_equal(out["depth"][2 * p, :nl], odp)
                assert dg == first[key], f"pair {i} (content {key}) differs from its first occurrence"
                total_matches += int((out["ur"][2 * p, :nl] >= 0).sum())
    assert seen == set(range(n_total))
    assert total_matches > 300 * n_total
    ex.close()


def test_cuda_path_equals_the_references_own_code(lib):
    """no oracle in between: the sm_100a path against oracle/_ref/libslam_ref.so = the reference's own orb_extractor.cpp +
    frame.cpp compiled unmodified (oracle/Makefile.ref; the prebuilt file travels to the GPU box).  Frame's stereo
    constructor there runs both extractions, UndistortKeyPoints and ComputeStereoMatches."""
    import reference_lib as R
    if not R.available():
        pytest.skip("oracle/_ref/libslam_ref.so not present")
    for seed in (3, 17):
        l, r = synth.stereo_pair(seed=seed)
        F = R.Frame(l, r)
        eL, eR = orbfe.ORBextractor(lib=lib), orbfe.ORBextractor(lib=lib)
        kl, dl = eL.Compute(l)
        kr, dr = eR.Compute(r)
        for f in P.KP_FIELDS:
            assert np.array_equal(kl[f], F.kps[f]) and np.array_equal(kr[f], F.kps_right[f]), f
        assert np.array_equal(dl, F.desc) and np.array_equal(dr, F.desc_right)
        bf, fx = np.float32(P.KITTI["bf"]), np.float32(P.KITTI["fx"])
        n, ur, dp = orbfe.ComputeStereoMatches(eL, eR, kl, dl, kr, dr, float(bf), float(bf / fx))
        assert n == int((F.u_right >= 0).sum()) and n > 500
        assert np.array_equal(ur, F.u_right) and np.array_equal(dp, F.depth)
        eL.close(); eR.close()
    # 8000 features on a 1080p frame (config 5) against ORBextractor::Compute itself
    img = synth.frame(1080, 1920, seed=3)
    ex = orbfe.ORBextractor(8000, lib=lib)
    k, d = ex.Compute(img)
    rk, rd = R.extract(img, 8000)
    for f in P.KP_FIELDS:
        assert np.array_equal(k[f], rk[f]), f
    assert np.array_equal(d, rd)
    ex.close()


def test_n2_search_local_points_device_resident(lib, kitti_two_frames):
    """Tracker::SearchLocalPoints (core/tracker.cpp:1196-1226) as one call: IsInFrustum chained into SearchByProjection in HBM"""
    ka, da, _, _, scale = kitti_two_frames
    rng = np.random.default_rng(81)
    ur = np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - rng.uniform(1, 60, len(ka)), -1).astype(np.float32)
    assert P.check_search_local_points(lib, ka, da, scale, 1241, 376, seed=82, u_right=ur, n_extra=18000) > 800
    assert P.check_search_local_points(lib, ka, da, scale, 1241, 376, seed=83, th=5, n_extra=3000) > 800


@pytest.mark.parametrize("nf,params", [(2000, (1.2, 1, 20, 7)), (3000, (1.1, 12, 20, 7)), (1500, (1.5, 5, 20, 7)), (1000, (2.0, 4, 20, 7)),
                                       (800, (2.5, 3, 20, 7)), (500, (1.3, 6, 12, 5)), (2000, (1.2, 8, 7, 7)), (100, (1.2, 8, 40, 12))])
def test_extractor_parameter_sweep(lib, nf, params):
    """the extractor is not KITTI-specific: 1 and 12 levels, scale 1.1 / 1.5, exact 2x (OpenCV's INTER_AREA path), > 2x (generic
    resize kernel), other thresholds (ini == min: no fallback round), small quotas -- stage by stage against the oracle"""
    P.check_extract(lib, synth.frame(seed=int(params[0] * 10) + params[1]), nfeatures=nf, params=params)


@pytest.mark.parametrize("nf,params,bf", [(1500, (1.5, 5, 20, 7), 386.1448), (1000, (2.0, 4, 20, 7), 386.1448), (3000, (1.1, 12, 20, 7), 120.0),
                                          (2000, (1.2, 1, 20, 7), 386.1448)])
def test_stereo_parameter_sweep(lib, nf, params, bf):
    """ComputeStereoMatches on other pyramids (the SAD refinement runs on the keypoint's own level) and another baseline"""
    l, r = synth.stereo_pair(seed=90 + params[1])
    assert P.check_stereo(lib, l, r, nfeatures=nf, bf=bf, params=params) > 50


def test_empty_inputs_everywhere(lib, kitti_two_frames):
    ka, da, _, _, scale = kitti_two_frames
    P.check_empty_inputs(lib, ka, da, scale)


def test_frame_from_extractor_device_resident(lib):
    """extract -> matcher view -> searches with the frame's own features never leaving the device"""
    l, r = synth.stereo_pair(seed=91)
    assert P.check_frame_from_extractor(lib, l, r, seed=92) > 500


@pytest.mark.parametrize("chunk", range(4))
def test_fuzz_slice_tie_heavy_content_and_random_parameters(lib, chunk):
    """a bounded slice of tests/fuzz_parity.py (the full sweep is run by hand: 1000+ cases identical)"""
    import fuzz_parity
    for seed in range(5000 + 30 * chunk, 5000 + 30 * (chunk + 1)):
        fuzz_parity.run_case(lib, seed)


def test_fuzz_slice_matchers_on_tie_heavy_keypoints():
    """tests/fuzz_matchers.py on the GPU build: every search routine on keypoints with identical descriptors everywhere"""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "fuzz_matchers.py"), "gpu", "9"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "failures 0" in r.stdout


def test_single_process_multi_gpu_runner_matches_per_rank_digests(lib):
    """shard.SequenceRunner (ONE process, a host thread and two streams per device) against the per-rank run of the same
    sequence: identical digests pair by pair; on a multi-GPU box every visible device takes a shard."""
    from slam_framework_b200 import shard
    n_pairs, bf, bl = 23, 386.1448, 386.1448 / 718.856
    pairs = [synth.stereo_pair(188, 620, seed=900 + i) for i in range(n_pairs)]
    ex = orbfe.ORBextractor(1000, lib=lib, max_images=2)

    def one(i):
        ex.upload(list(pairs[i])); ex.run(2); ex.run_stereo(1, bf, bl)
        b = ex.download(2, ex.make_buffers(2, stereo=True))
        n0, n1 = b["n"]
        return [b["kps"][0, :n0], b["desc"][0, :n0], b["kps"][1, :n1], b["desc"][1, :n1], b["ur"][0, :n0], b["depth"][0, :n0]]
    single = shard.process_sequence(one, n_pairs, 0, 1)
    ex.close()
    ndev = lib.orbfe_device_count()
    for devices in ([0], list(range(ndev)) if ndev > 1 else [0, 0, 0]):
        runner = shard.SequenceRunner(lib, devices=devices, params=dict(nfeatures=1000), batch_pairs=4, lanes=2)
        assert runner.run(lambda i: pairs[i], n_pairs, bf, bl) == single
        frames = np.stack([im for p in pairs for im in p])
        assert runner.run(None, n_pairs, bf, bl, get_batch=lambda s, e: frames[2 * s:2 * e]) == single   # the batch form, kept handles
        runner.close()


def test_extract_graph_replay_equals_plain_path(lib, monkeypatch):
    """orbfe_extract replays a captured CUDA graph from the second frame of a geometry on; the results must equal the
    stream-ordered path (ORBFE_NO_GRAPH) bit for bit, for one frame and for a two-frame batch, across changing content."""
    imgs = [synth.frame(seed=70 + i) for i in range(4)]
    ex = orbfe.ORBextractor(2000, lib=lib, max_images=2)
    got = [ex.Compute(im) for im in imgs] + [ex.Compute(imgs[0])]
    got2 = ex.extract_batch([imgs[1], imgs[2]])
    monkeypatch.setenv("ORBFE_NO_GRAPH", "1")
    ref = orbfe.ORBextractor(2000, lib=lib, max_images=2)
    want = [ref.Compute(im) for im in imgs] + [ref.Compute(imgs[0])]
    want2 = ref.extract_batch([imgs[1], imgs[2]])
    for (k, d), (rk, rd) in zip(got + got2, want + want2):
        assert np.array_equal(k, rk) and np.array_equal(d, rd)
    ex.close(); ref.close()


def test_n2_distorted_camera_frame_equals_the_references_own_code(lib):
    """the k1 != 0 path (frame.cpp:614-673) without the oracle in between: the reference's own stereo Frame constructor on a
    distorted camera against orbfe_undistort_keypoints (keypoints and the four image corners = ComputeImageBounds), the stereo
    matches on the original keypoints, and the grid searches of a frame handle built with those bounds."""
    import reference_lib as R
    if not R.available():
        pytest.skip("oracle/_ref/libslam_ref.so not present")
    dist = np.array([-0.28, 0.07, 0.0002, -0.0002], np.float32)
    fx, fy, cx, cy, bf = 718.856, 718.856, 607.1928, 185.2157, 386.1448   # reference_lib.Frame's defaults
    l, r = synth.stereo_pair(seed=17)
    try:
        R.set_test_distortion(dist)
        F = R.Frame(l, r)
    finally:
        R.set_test_distortion(None)
    eL, eR = orbfe.ORBextractor(lib=lib), orbfe.ORBextractor(lib=lib)
    kl, dl = eL.Compute(l)
    kr, dr = eR.Compute(r)
    un = orbfe.UndistortKeyPoints(kl, fx, fy, cx, cy, dist, lib=lib)
    for f in P.KP_FIELDS:
        assert np.array_equal(un[f], F.kps_un[f]), f
    h, w = l.shape
    corners = np.zeros(4, orbfe.KP_DTYPE)
    corners["x"], corners["y"] = [0, w, 0, w], [0, 0, h, h]
    c = orbfe.UndistortKeyPoints(corners, fx, fy, cx, cy, dist, lib=lib)
    bounds = (min(c["x"][0], c["x"][2]), max(c["x"][1], c["x"][3]), min(c["y"][0], c["y"][1]), max(c["y"][2], c["y"][3]))
    assert np.array_equal(np.array(bounds, np.float32), F.bounds)
    n, ur, dp = orbfe.ComputeStereoMatches(eL, eR, kl, dl, kr, dr, bf, float(np.float32(bf) / np.float32(fx)))
    assert np.array_equal(ur, F.u_right) and np.array_equal(dp, F.depth) and n > 100
    G = orbfe.Frame(un, dl, eL.GetScaleFactors(), bounds, ur, lib=lib)
    rng = np.random.default_rng(2)
    for _ in range(40):
        x, y, rad = float(rng.uniform(bounds[0], bounds[1])), float(rng.uniform(bounds[2], bounds[3])), float(rng.uniform(5, 120))
        assert np.array_equal(G.GetFeaturesInArea(x, y, rad), F.features_in_area(x, y, rad))
    G.close(); eL.close(); eR.close()
