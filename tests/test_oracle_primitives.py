"""Pins the CPU oracle's OpenCV primitives against golden vectors produced by cv2 4.13
(tests/golden/make_golden.py) and, when cv2 is importable, against live cv2 on extra cases."""
import numpy as np
import pytest

import oracle_lib as O
from slam_framework_b200 import synth

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_resize_golden(golden):
    for case in golden["resize_cases"]:
        src, dst = golden[f"{case}_src"], golden[f"{case}_dst"]
        got = O.resize_linear(src, dst.shape[1], dst.shape[0])
        assert np.array_equal(got, dst), case


def test_border_golden(golden):
    got = O.border_reflect101(golden["border_src"], 19)
    assert np.array_equal(got, golden["border_dst"])


def test_gaussian_golden(golden):
    for case in golden["gauss_cases"]:
        assert np.array_equal(O.gaussian7x7(golden[f"{case}_src"]), golden[f"{case}_dst"]), case


def test_fast_golden(golden):
    for case in golden["fast_cases"]:
        _, i, th = case.split("_")
        ref = golden[f"{case}_kps"]
        got = O.fast9(golden[f"fast_{i}_src"], int(th), True)
        assert len(got) == len(ref), case
        if len(ref):
            assert np.array_equal(got["x"], ref[:, 0]) and np.array_equal(got["y"], ref[:, 1]), case
            assert np.array_equal(got["response"], ref[:, 4]), case
            assert np.all(got["size"] == 7) and np.all(got["angle"] == -1) and np.all(got["octave"] == 0)


def test_fast_atan2_golden(golden):
    got = np.array([O.fast_atan2(y, x) for y, x in zip(golden["atan_y"], golden["atan_x"])], np.float32)
    assert np.array_equal(got.view(np.uint32), golden["atan_deg"].view(np.uint32))


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.integers(0, 256, 32, dtype=np.uint8)
        b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())


def test_tables_kitti():
    t = O.Extractor(2000, 1.2, 8, 20, 7).tables()
    assert list(t["features_per_level"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(O.Extractor(4000, 1.2, 8, 20, 7).tables()["features_per_level"]) == [869, 724, 603, 503, 419, 349, 291, 242]
    assert list(O.Extractor(8000, 1.2, 8, 20, 7).tables()["features_per_level"]) == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    s = t["scale"]
    exp = [np.float32(1.0)]
    for _ in range(7):
        exp.append(np.float32(np.float64(exp[-1]) * np.float64(np.float32(1.2))))
    assert np.array_equal(s, np.array(exp, np.float32))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_cv2_pyramid_blur_kitti():
    img = synth.frame(seed=5)
    ex = O.Extractor()
    ex.extract(img)
    prev = img
    for l in range(1, 8):
        p = ex.pyramid_level(l)
        ref = cv2.resize(prev, (p.shape[1], p.shape[0]), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, p)
        assert np.array_equal(cv2.copyMakeBorder(ref, 19, 19, 19, 19, cv2.BORDER_REFLECT_101), ex.pyramid_padded(l))
        assert np.array_equal(cv2.GaussianBlur(ref, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101), ex.blurred(l))
        prev = ref


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_cv2_resize_sizes():
    rng = np.random.default_rng(3)
    for _ in range(25):
        sh, sw = int(rng.integers(8, 200)), int(rng.integers(8, 300))
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        f = rng.uniform(1.05, 2.6)
        dw, dh = max(int(round(sw / f)), 2), max(int(round(sh / f)), 2)
        assert np.array_equal(O.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    img = rng.integers(0, 256, (96, 130), dtype=np.uint8)  # exact 2x -> OpenCV runs INTER_AREA
    assert np.array_equal(O.resize_linear(img, 65, 48), cv2.resize(img, (65, 48), interpolation=cv2.INTER_LINEAR))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_cv2_percell_fast_equals_oracle_candidates():
    """Grid loop of ComputeKeyPointsOctTree (orb_extractor.cpp:706-770) replayed with cv2.FAST per
    cell must reproduce the oracle's candidate list, order included."""
    img = synth.frame(seed=9)
    ex = O.Extractor()
    ex.extract(img)
    for level in (0, 3, 7):
        pl = ex.pyramid_level(level)
        h, w = pl.shape
        minB, maxBX, maxBY = 16, w - 16, h - 16
        width, height = float(maxBX - minB), float(maxBY - minB)
        nCols, nRows = int(width / 30), int(height / 30)
        wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
        out = []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = min(iniY + hCell + 6, maxBY)
            if iniY >= maxBY - 3:
                continue
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = min(iniX + wCell + 6, maxBX)
                if iniX >= maxBX - 6:
                    continue
                cell = np.ascontiguousarray(pl[iniY:maxY, iniX:maxX])
                kps = cv2.FastFeatureDetector_create(20, True).detect(cell, None)
                if not kps:
                    kps = cv2.FastFeatureDetector_create(7, True).detect(cell, None)
                out += [(k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response) for k in kps]
        c = ex.candidates(level)
        assert len(c) == len(out)
        ref = np.array(out, np.float32).reshape(-1, 3)
        assert np.array_equal(c["x"], ref[:, 0]) and np.array_equal(c["y"], ref[:, 1])
        assert np.array_equal(c["response"], ref[:, 2])


def test_cvt_gray_golden(golden):
    assert np.array_equal(O.cvt_gray(golden["gray_src3"], True), golden["gray_rgb"])
    assert np.array_equal(O.cvt_gray(golden["gray_src3"], False), golden["gray_bgr"])
    assert np.array_equal(O.cvt_gray(golden["gray_src4"], True), golden["gray_rgba"])
    assert np.array_equal(O.cvt_gray(golden["gray_src4"], False), golden["gray_bgra"])


# ---- N2: the OpenCV arithmetic of the Frame tail (tests/golden/make_golden_frame_tail.py) -----------------------
@pytest.fixture(scope="module")
def golden_tail():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_frame_tail.npz"))


def test_undistort_points_golden(golden_tail):
    g = golden_tail
    for c in range(int(g["und_cases"])):
        K, d = g[f"und_{c}_K"], g[f"und_{c}_dist"]
        got = O.undistort_points(g[f"und_{c}_src"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], d)
        assert np.array_equal(got, g[f"und_{c}_dst"]), f"camera {c}: {(got != g[f'und_{c}_dst']).any(1).sum()} points differ"
    # Frame::UndistortKeyPoints short-circuit (frame.cpp:616-619): dist[0] == 0 => unchanged
    src = g["und_0_src"]
    assert np.array_equal(O.undistort_points(src, 700.0, 700.0, 600.0, 180.0, np.zeros(4, np.float32)), src)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_undistort_points_live_cv2():
    rng = np.random.default_rng(77)
    for _ in range(5):
        fx, fy = rng.uniform(300, 900, 2)
        cx, cy = rng.uniform(200, 700), rng.uniform(100, 400)
        d = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01),
                      rng.uniform(-0.1, 0.1)], np.float32)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        pts = np.stack([rng.uniform(0, 1241, 2000), rng.uniform(0, 376, 2000)], 1).astype(np.float32)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
        assert np.array_equal(O.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], d), ref)


def test_frustum_gemm_and_norm_golden(golden_tail):
    """Rcw*P + tcw and cv::norm(PO) as Frame::IsInFrustum evaluates them: the oracle's projection must reproduce cv2's
    gemm bit for bit (checked through proj = PcX/PcZ with fx = 1, cx = 0) and its distance cv2's norm (through the
    min/max distance gate)."""
    g = golden_tail
    R, P, t, ref = g["gemm_R"], g["gemm_P"], g["gemm_t"], g["gemm_out"]
    n = len(R)
    ok = 0
    for i in range(n):
        Pc = ref[i, :, 0]
        if Pc[2] <= 0:
            continue
        w = P[i, :, 0][None, :]
        nrm = np.array([[0, 0, 1]], np.float32)
        big = (-1e30, 1e30, -1e30, 1e30)
        _, out = O.is_in_frustum(w, nrm, [0.0], [1e30], [1e30], R[i], t[i, :, 0], np.zeros(3, np.float32), 1.0, 1.0, 0.0, 0.0, 0.0, big,
                                 np.float32(np.log(np.float32(1.2))), 8, -2.0)
        invz = np.float32(1.0) / Pc[2]
        assert out["in_view"][0] == 1
        assert out["proj_x"][0] == np.float32(np.float32(np.float32(1.0) * Pc[0]) * invz) + np.float32(0.0), i
        assert out["proj_y"][0] == np.float32(np.float32(np.float32(1.0) * Pc[1]) * invz) + np.float32(0.0), i
        # distance gate straddling cv2's norm: dist = (float)norm must pass [dist, dist] and fail (dist, inf)
        dist = np.float32(g["norm_out"][i])
        c1, _ = O.is_in_frustum(w, nrm, [dist], [dist], [dist], R[i], t[i, :, 0], np.zeros(3, np.float32), 1.0, 1.0, 0.0, 0.0, 0.0, big,
                                np.float32(0.18), 8, -2.0)
        c2, _ = O.is_in_frustum(w, nrm, [np.nextafter(dist, np.float32(np.inf))], [1e30], [1e30], R[i], t[i, :, 0], np.zeros(3, np.float32),
                                1.0, 1.0, 0.0, 0.0, 0.0, big, np.float32(0.18), 8, -2.0)
        assert c1 == 1 and c2 == 0, i
        ok += 1
    assert ok > 500
