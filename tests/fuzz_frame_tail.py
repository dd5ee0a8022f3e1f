#!/usr/bin/env python
"""Seed sweep of the Frame-tail rows (UndistortKeyPoints, IsInFrustum + PredictScale, the device-resident
Frame built from an extractor slot, SearchLocalPoints, the restated logf) against the oracle.
    python tests/fuzz_frame_tail.py [emu|gpu] [n_seeds]"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import oracle_lib as O, parity_common as P
from slam_framework_b200 import orbfe, synth

if len(sys.argv) < 2 or sys.argv[1] == "emu":
    from emu import build_emu
    L = orbfe.load(build_emu.build(), _test_emulation=True)
else:
    L = orbfe.load()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bad = 0
for seed in range(n):
    h, w = (200, 640) if seed % 2 else (376, 1241)
    kps, desc = O.Extractor(1000).extract(synth.frame(h, w, seed=100 + seed))
    scale = O.Extractor(1000).tables()["scale"]
    checks = [("undistort", lambda: P.check_undistort_keypoints(L, kps, seed=seed)),
              ("frustum", lambda: P.check_is_in_frustum(L, 4000, seed=seed)),
              ("logf", lambda: P.check_logf(L, 50000, seed=seed)),
              ("local_points", lambda: P.check_search_local_points(L, kps, desc, scale, w, h, seed=seed, n_extra=1500))]
    if seed % 3 == 0:
        l, r = synth.stereo_pair(h, w, seed=200 + seed)
        checks.append(("frame_from_extractor", lambda: P.check_frame_from_extractor(L, l, r, nfeatures=800, seed=seed)))
    for name, fn in checks:
        try:
            fn()
        except AssertionError as e:
            bad += 1
            print("FAIL seed", seed, name, str(e)[:150], flush=True)
        except Exception as e:
            bad += 1
            print("ERROR seed", seed, name, type(e).__name__, str(e)[:150], flush=True)
print("fuzz_frame_tail: failures", bad)
sys.exit(1 if bad else 0)
