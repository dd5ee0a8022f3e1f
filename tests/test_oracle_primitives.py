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
