#!/usr/bin/env python
"""Diagnostic: where a multi-threaded single-process driver spends its wall clock.  N threads, one device each, 60 batches of 64
pairs; prints the mean duration of every C-ABI call per thread."""
import ctypes, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench as B
from slam_framework_b200 import orbfe
n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 2
same_dev = len(sys.argv) > 2 and sys.argv[2] == "same"
L = orbfe.load()
pairs = B.make_pairs(8, 0)
frames = orbfe.pinned_empty((128, B.H, B.W), np.uint8, lib=L)
for p in range(64):
    frames[2 * p], frames[2 * p + 1] = pairs[p % 8]
ptrs = (ctypes.c_void_p * 128)(*[frames[i].ctypes.data for i in range(128)])
res = {}


def work(t):
    dev = 0 if same_dev else t
    lanes = []
    for _ in range(3):
        ex = orbfe.ORBextractor(B.NFEATURES, B.SCALE, B.NLEVELS, B.INI_TH, B.MIN_TH, device=dev, max_images=128, max_size=(B.W, B.H), lib=L)
        cap = ex.max_keypoints()
        buf = dict(kps=orbfe.pinned_empty((128, cap), orbfe.KP_DTYPE, lib=L), desc=orbfe.pinned_empty((128, cap, 32), np.uint8, lib=L),
                   n=orbfe.pinned_empty((128,), np.int32, lib=L), cap=cap, ur=orbfe.pinned_empty((128, cap), np.float32, lib=L),
                   depth=orbfe.pinned_empty((128, cap), np.float32, lib=L))
        lanes.append((ex, buf))
    acc = {k: 0.0 for k in ("sync", "upload", "run", "stereo", "download")}
    for it in range(70):
        ex, buf = lanes[it % 3]
        t0 = time.perf_counter(); ex.sync(); t1 = time.perf_counter()
        ex.upload_ptrs(ptrs, 128, B.W, B.H, B.W); t2 = time.perf_counter()
        ex.run(128); t3 = time.perf_counter()
        ex.run_stereo(64, B.BF, B.BF / B.FX); t4 = time.perf_counter()
        ex.download_async(128, buf); t5 = time.perf_counter()
        if it >= 10:
            for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                acc[k] += v
    for ex, _ in lanes:
        ex.sync()
    res[t] = {k: round(1e3 * v / 60, 3) for k, v in acc.items()}


t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
[t.start() for t in th]; [t.join() for t in th]
print(json.dumps({"threads": n_threads, "same_device": same_dev, "wall_s": round(time.perf_counter() - t0, 3), "ms_per_call": res}))
