#!/usr/bin/env python
"""Every OrbMatcher routine (and the vocabulary transform, SearchLocalPoints) against the oracle on keypoints extracted from
tie-heavy images -- checkerboards, rectangles, gratings: many identical descriptors, so every best / second-best / occupancy
tie rule of the searches is exercised -- and on ordinary texture, with per-seed thresholds, ratios, windows and vocabulary shapes.      python tests/fuzz_matchers.py [emu|gpu] [n_seeds]"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import oracle_lib as O, parity_common as P, fuzz_parity as F
from emu import build_emu
from slam_framework_b200 import orbfe
L = orbfe.load(build_emu.build(), _test_emulation=True) if len(sys.argv) < 2 or sys.argv[1] == "emu" else orbfe.load()
scale = O.Extractor(1500).tables()["scale"]
bad = 0
for seed in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    rng = np.random.default_rng(seed)
    kind = ("checker", "rects", "grating", "texture", "mix")[seed % 5]
    pick = lambda *v: v[int(rng.integers(len(v)))]  # per-seed thresholds / ratios / windows
    h, w = 200, 640
    img = np.clip(np.rint(F.content(kind, h, w + 12, rng)), 0, 255).astype(np.uint8)
    a, b = np.ascontiguousarray(img[:, 8:w + 8]), np.ascontiguousarray(img[4:, :w])
    b = np.vstack([b, b[-4:]])
    ka, da = O.Extractor(1000).extract(a)
    kb, db = O.Extractor(1000).extract(b)
    if len(ka) < 50 or len(kb) < 50:
        print("seed", seed, kind, "too few keypoints", len(ka), len(kb)); continue
    ur = np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - rng.uniform(1, 60, len(ka)), -1).astype(np.float32)
    checks = [
        ("projection_mappoints", lambda: P.check_search_by_projection_mappoints(L, ka, da, scale, w, h, 3000, seed=5 + seed, u_right=ur, th=pick(1, 3, 5), nnratio=pick(0.6, 0.8, 0.9))),
        ("projection_lastframe", lambda: P.check_search_by_projection_lastframe(L, ka, da, scale, w, h, seed=7 + seed, u_right=ur, th=pick(3.0, 7.0, 15.0))),
        ("initialization", lambda: P.check_search_for_initialization(L, a, b, lambda im, nf: O.Extractor(nf).extract(im), nfeatures=1000, window=pick(50, 100, 200))),
        ("bow", lambda: P.check_search_by_bow(L, kb, db, ka, da, scale, w, h, seed=3 + seed, nnratio=pick(0.6, 0.7, 0.9))),
        ("projection_sim3", lambda: P.check_search_by_projection_sim3(L, ka, da, scale, w, h, 2000, seed=11 + seed, th=pick(5, 10, 20))),
        ("projection_keyframe", lambda: P.check_search_by_projection_keyframe(L, ka, da, scale, w, h, 1200, seed=12 + seed, th=pick(5.0, 10.0), orb_dist=pick(50, 100))),
        ("fuse", lambda: P.check_fuse(L, ka, da, scale, w, h, 2000, seed=13 + seed, u_right=ur, th=pick(2.0, 3.0, 4.0))),
        ("sim3", lambda: P.check_search_by_sim3(L, ka, da, kb, db, scale, w, h, seed=14 + seed, shift=(8.0, 4.0), th=pick(5.0, 7.5, 10.0))),
        ("bow_keyframes", lambda: P.check_search_by_bow_keyframes(L, ka, da, kb, db, scale, w, h, seed=15 + seed, nnratio=pick(0.6, 0.8, 0.9))),
        ("triangulation", lambda: P.check_search_for_triangulation(L, ka, da, kb, db, scale, w, h, seed=16 + seed)),
        ("bow_transform", lambda: P.check_bow_transform(L, da, seed=21 + seed, k=pick(4, 7, 10), L=pick(3, 4, 5))),
        ("local_points", lambda: P.check_search_local_points(L, ka, da, scale, w, h, seed=42 + seed, n_extra=700, th=pick(1, 3))),
    ]
    for name, fn in checks:
        try:
            fn()
        except AssertionError as e:
            bad += 1
            print("FAIL seed", seed, kind, name, str(e)[:150], flush=True)
        except Exception as e:
            bad += 1
            print("ERROR seed", seed, kind, name, type(e).__name__, str(e)[:150], flush=True)
    print("seed", seed, kind, len(ka), len(kb), "done", flush=True)
print("fuzz_matchers: failures", bad)
sys.exit(1 if bad else 0)
