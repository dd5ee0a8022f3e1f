#!/usr/bin/env python
"""A/B probe for k_match_solve build variants (ORBFE_LIB=...): wall clock of SearchByProjection(F, 20 000 map points) on a 1080p
frame with 8000 features and of SearchByProjection(Cur, Last) on a KITTI frame; prints one JSON line."""
import json, os, statistics, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as P
from slam_framework_b200 import orbfe, synth


def med(f, reps=200):
    for _ in range(10):
        f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(t)


L = orbfe.load()
img = synth.frame(1080, 1920, seed=1920)
ex = orbfe.ORBextractor(8000, lib=L)
kps, desc = ex.Compute(img)
F = orbfe.Frame(kps, desc, ex.GetScaleFactors(), (0, 1920, 0, 1080), lib=L)
mp = P.synth_map_points(kps, desc, np.random.default_rng(3), 20000)
args = (mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"], mp["desc"], mp["has_obs"], mp["occupied"])
m = orbfe.OrbMatcher(0.8)
out = {"lib": os.environ.get("ORBFE_LIB", "default"), "sbp_20k_ms": med(lambda: m.SearchByProjectionMapPoints(F, *args, 1))}
l, r = synth.stereo_pair(seed=31)
ex2 = orbfe.ORBextractor(lib=L)
kl, dl = ex2.Compute(l)
F2 = orbfe.Frame(kl, dl, ex2.GetScaleFactors(), (0, 1241, 0, 376), lib=L)
rng = np.random.default_rng(8); nk = len(kl)
u = (kl["x"] + rng.uniform(-5, 5, nk)).astype(np.float32); v = (kl["y"] + rng.uniform(-5, 5, nk)).astype(np.float32)
iz = rng.uniform(0.01, 0.2, nk).astype(np.float32); octv = kl["octave"].astype(np.int32); ang = kl["angle"].copy()
la = (np.ones(nk, np.uint8), u, v, iz, octv, ang, dl, np.ones(nk, np.uint8), P.KITTI["bf"], 0, 0, np.zeros(nk, np.uint8), 7.0)
m2 = orbfe.OrbMatcher(0.9, True)
out["lastframe_2k_ms"] = med(lambda: m2.SearchByProjectionLastFrame(F2, *la))
print(json.dumps(out))
