#!/usr/bin/env python
"""Profiling driver: ONE batch of `--pairs` distinct synthetic KITTI-shaped stereo pairs (the batch bench.py times) through
orbfe_run + orbfe_run_stereo, bracketed by cudaProfilerStart/Stop so that ncu captures exactly that step:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/profile_step.py --pairs 64
    ncu --profile-from-start off --set full --clock-control none --import-source on -o full python tools/profile_step.py --pairs 64

Without ncu it prints the per-stage CUDA-event times of the same step (not a bench value)."""
import argparse, ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (constants and the synthetic-pair maker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    import torch
    from slam_framework_b200 import orbfe
    L = orbfe.load()
    n_img = 2 * a.pairs
    ex = orbfe.ORBextractor(B.NFEATURES, B.SCALE, B.NLEVELS, B.INI_TH, B.MIN_TH, device=0, max_images=n_img, max_size=(B.W, B.H), lib=L)
    pairs = B.make_pairs(a.pairs, 0)
    imgs = [im for p in pairs for im in p]
    ex.upload(imgs)
    for _ in range(a.warm):
        ex.run(n_img); ex.run_stereo(a.pairs, B.BF, B.BF / B.FX)
    ex.sync()
    ex.set_stage_timing(True); ex.stage_summary()
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    for _ in range(a.steps):
        ex.run(n_img); ex.run_stereo(a.pairs, B.BF, B.BF / B.FX)
    ex.sync()
    rt.cudaProfilerStop()
    st, runs = ex.stage_summary()
    print(json.dumps({"pairs": a.pairs, "stage_ms_per_batch": {k: v / max(runs, 1) for k, v in st.items()}}))
    ex.close()


if __name__ == "__main__":
    main()
