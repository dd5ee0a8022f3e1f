#!/usr/bin/env python
"""Generates tests/golden/cv2_primitives.npz: known-answer vectors for the OpenCV primitives on
the hot path, produced by the Python cv2 build of the same library (the reference's arithmetic
for resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 lives in un-vendored OpenCV).
The reference itself ships no golden vectors, so these pin the oracle.

Run in the build container:  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from slam_framework_b200 import synth  # noqa: E402


def main():
    cv2.setNumThreads(1)
    rng = np.random.default_rng(1234)
    out = {"cv2_version": np.array(cv2.__version__)}
    # --- resize chain on a KITTI-shaped crop and odd sizes
    cases = []
    base = synth.frame(120, 403, seed=7)
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    for name, img, sizes in (("synth", base, [(336, 100), (280, 83), (233, 69), (201, 60)]),
                             ("noise", noise, [(109, 81), (91, 67), (65, 48), (33, 25)]),
                             ("half", noise[:96, :130], [(65, 48)])):
        cur = img
        for i, (dw, dh) in enumerate(sizes):
            dst = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
            out[f"resize_{name}_{i}_src"] = cur
            out[f"resize_{name}_{i}_dst"] = dst
            cases.append(f"resize_{name}_{i}")
            cur = dst
    out["resize_cases"] = np.array(cases)
    # --- border
    out["border_src"] = noise[:40, :57].copy()
    out["border_dst"] = cv2.copyMakeBorder(out["border_src"], 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    # --- gaussian
    gcases = []
    for i, img in enumerate((base, noise, noise[:9, :11].copy(), synth.frame(64, 64, seed=3))):
        out[f"gauss_{i}_src"] = img
        out[f"gauss_{i}_dst"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        gcases.append(f"gauss_{i}")
    out["gauss_cases"] = np.array(gcases)
    # --- FAST (TYPE_9_16, nms on), incl. tiny cell-sized images
    fcases = []
    imgs = [base, noise, synth.frame(38, 37, seed=11), synth.frame(46, 38, seed=12), noise[:7, :7].copy(),
            noise[:8, :30].copy()]
    for i, img in enumerate(imgs):
        for th in (7, 20, 40):
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                 type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kps = det.detect(img, None)
            arr = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps],
                           np.float32).reshape(-1, 7)
            out[f"fast_{i}_{th}_kps"] = arr
            fcases.append(f"fast_{i}_{th}")
        out[f"fast_{i}_src"] = img
    out["fast_cases"] = np.array(fcases)
    # --- fastAtan2
    ys = np.concatenate([rng.integers(-200000, 200000, 4000), [0, 0, 1, -1, 0, 5, -5]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-200000, 200000, 4000), [0, 1, 0, 0, -1, 5, -5]]).astype(np.float32)
    out["atan_y"] = ys
    out["atan_x"] = xs
    out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    # --- cvtColor to gray (tracker.cpp:110-127)
    col3 = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    col4 = rng.integers(0, 256, (21, 50, 4), dtype=np.uint8)
    out["gray_src3"], out["gray_src4"] = col3, col4
    out["gray_rgb"] = cv2.cvtColor(col3, cv2.COLOR_RGB2GRAY)
    out["gray_bgr"] = cv2.cvtColor(col3, cv2.COLOR_BGR2GRAY)
    out["gray_rgba"] = cv2.cvtColor(col4, cv2.COLOR_RGBA2GRAY)
    out["gray_bgra"] = cv2.cvtColor(col4, cv2.COLOR_BGRA2GRAY)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)
    print("wrote", os.path.join(HERE, "cv2_primitives.npz"))


if __name__ == "__main__":
    main()
