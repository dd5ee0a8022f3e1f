#!/usr/bin/env python
"""TEST-SIDE measurement tool (it loads the oracle as checker and CPU arm, so it lives under tests/).  Secondary measurement: the matcher rows (M2-M4) of SURVEY.md 8 on the GPU next to the CPU oracle, on the
BASELINE config-4 / config-5 shapes.  Wall clock of the C-ABI call (host arrays in, host arrays out), median of
`reps` calls; prints one JSON line.  usage: python tests/bench_matchers.py [--reps 20]"""
import argparse, json, os, statistics, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_lib as O
import parity_common as P
from slam_framework_b200 import orbfe, synth


def med(f, reps):
    f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(t)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=20); a = ap.parse_args()
    L = orbfe.load()
    out = {}
    # config 5: 1080p, 8000 features, 20k projected map points
    for (h, w) in ((1080, 1920), (2160, 3840)):
        img = synth.frame(h, w, seed=w)
        ex = orbfe.ORBextractor(8000, lib=L)
        kps, desc = ex.Compute(img)
        t_ext = med(lambda: ex.Compute(img), 5)
        oe = O.Extractor(8000)
        t0 = time.perf_counter(); oe.extract(img); t_ext_cpu = (time.perf_counter() - t0) * 1e3
        scale = ex.GetScaleFactors()
        F, OF = P.make_frames(kps, desc, scale, w, h, L)
        mp = P.synth_map_points(kps, desc, np.random.default_rng(3), 20000)
        args = (mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"], mp["desc"], mp["has_obs"], mp["occupied"])
        m = orbfe.OrbMatcher(0.8)
        n, asg = m.SearchByProjectionMapPoints(F, *args, 1)
        on, oasg = O.search_by_projection_mappoints(OF, *args, 1, 0.8)
        assert n == on and np.array_equal(asg, oasg)
        out[f"config5_{w}x{h}"] = {
            "extract_8000_ms_gpu": t_ext, "extract_8000_ms_cpu_oracle_1thread": t_ext_cpu, "keypoints": len(kps),
            "search_by_projection_20k_ms_gpu": med(lambda: m.SearchByProjectionMapPoints(F, *args, 1), a.reps),
            "search_by_projection_20k_ms_cpu_oracle_1thread": med(lambda: O.search_by_projection_mappoints(OF, *args, 1, 0.8), 5),
            "matches": n}
        ex.close()
    # config 4: mono, 4000 features, SearchForInitialization
    a_img, b_img = synth.shifted_frame(21, dx=8, dy=4)
    ex = orbfe.ORBextractor(4000, lib=L)
    k1, d1 = ex.Compute(a_img); k2, d2 = ex.Compute(b_img)
    scale = ex.GetScaleFactors()
    F1, OF1 = P.make_frames(k1, d1, scale, 1241, 376, L); F2, OF2 = P.make_frames(k2, d2, scale, 1241, 376, L)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    m = orbfe.OrbMatcher(0.9, True)
    n, m12, _ = m.SearchForInitialization(F1, F2, prev, 100)
    on, om12, _ = O.search_for_initialization(OF1, OF2, prev, 100, 0.9, True)
    assert n == on and np.array_equal(m12, om12)
    out["config4_mono_4000"] = {
        "extract_4000_ms_gpu": med(lambda: ex.Compute(a_img), 10),
        "search_for_initialization_ms_gpu": med(lambda: m.SearchForInitialization(F1, F2, prev, 100), a.reps),
        "search_for_initialization_ms_cpu_oracle_1thread": med(lambda: O.search_for_initialization(OF1, OF2, prev, 100, 0.9, True), 5),
        "matches": n}
    # tracking: SearchByProjection(Cur, Last) on a KITTI frame
    l, r = synth.stereo_pair(seed=31)
    ex2 = orbfe.ORBextractor(lib=L)
    kl, dl = ex2.Compute(l)
    F, OF = P.make_frames(kl, dl, ex2.GetScaleFactors(), 1241, 376, L)
    rng = np.random.default_rng(8); nk = len(kl)
    u = (kl["x"] + rng.uniform(-5, 5, nk)).astype(np.float32); v = (kl["y"] + rng.uniform(-5, 5, nk)).astype(np.float32)
    iz = rng.uniform(0.01, 0.2, nk).astype(np.float32); octv = kl["octave"].astype(np.int32); ang = kl["angle"].copy()
    valid = np.ones(nk, np.uint8); has = np.ones(nk, np.uint8); occ = np.zeros(nk, np.uint8)
    m = orbfe.OrbMatcher(0.9, True)
    la = (valid, u, v, iz, octv, ang, dl, has, P.KITTI["bf"], 0, 0, occ, 7.0)
    n, asg = m.SearchByProjectionLastFrame(F, *la)
    on, oasg = O.search_by_projection_lastframe(OF, *la, True)
    assert n == on and np.array_equal(asg, oasg)
    out["tracking_kitti_2000"] = {"search_by_projection_lastframe_ms_gpu": med(lambda: m.SearchByProjectionLastFrame(F, *la), a.reps),
                                  "search_by_projection_lastframe_ms_cpu_oracle_1thread": med(lambda: O.search_by_projection_lastframe(OF, *la, True), 5),
                                  "frame_create_ms_gpu": med(lambda: orbfe.Frame(kl, dl, ex2.GetScaleFactors(), (0, 1241, 0, 376), lib=L).close(), a.reps),
                                  "matches": n}
    # N1 / N3: the other searches and the vocabulary transform on KITTI-sized keyframes (2000 features)
    a_img, b_img = synth.shifted_frame(43, dx=6, dy=0)
    ka, da = ex2.Compute(a_img); kb, db = ex2.Compute(b_img)
    scale = ex2.GetScaleFactors()
    FA, OFA = P.make_frames(ka, da, scale, 1241, 376, L); FB, OFB = P.make_frames(kb, db, scale, 1241, 376, L)
    rng = np.random.default_rng(5)
    mp = P._proj_points(ka, da, rng, 6000)
    ur = (mp["u"] - rng.uniform(0, 40, 6000)).astype(np.float32)
    fu = (mp["valid"], mp["u"], mp["v"], ur, mp["lvl"], mp["desc"], 3.0)
    assert np.array_equal(orbfe.Fuse(FA, *fu)[1], O.fuse(OFA, *fu)[1])
    matched_in = np.zeros(len(ka), np.uint8)
    sp = (mp["valid"], mp["u"], mp["v"], mp["lvl"], mp["desc"], matched_in, 10)
    assert np.array_equal(orbfe.SearchByProjectionSim3(FA, *sp)[1], O.search_by_projection_sim3(OFA, *sp)[1])
    varr = P.synth_vocabulary_uniform(np.random.default_rng(61), 10, 6, seed_desc=da[0].copy())
    V, OV = orbfe.OrbVocabulary(10, 6, 0, 0, *varr, lib=L), O.Vocabulary(10, 6, 0, 0, *varr)
    ta, tb = V.transform(da, 4), V.transform(db, 4)
    P.assert_bow_equal(ta, OV.transform(da, 4))
    fva, fvb = orbfe.feature_vector_dict(ta["fv"]), orbfe.feature_vector_dict(tb["fv"])
    ones_a, ones_b = np.ones(len(ka), np.uint8), np.ones(len(kb), np.uint8)
    fva, fvb = orbfe.flatten_feature_vector(fva), orbfe.flatten_feature_vector(fvb)   # flat once, outside the timed calls (both arms)
    bk = (da, ka["angle"], ones_a, ones_b, fva, fvb, 0.8, True)
    assert np.array_equal(orbfe.SearchByBoWKeyFrames(FB, *bk)[1], O.search_by_bow_keyframes(OFB, *bk)[1])
    F12 = np.array([[0, 0, 0], [0, 0, 6.0], [0, -6.0, 0]], np.float32)
    tr = (ka, da, ones_a, np.zeros(len(ka), np.uint8), ones_b, fva, fvb, F12, 1e6, 188.0, False, True)
    assert np.array_equal(orbfe.SearchForTriangulation(FB, *tr)[1], O.search_for_triangulation(OFB, *tr)[1])
    out["n1_n3_kitti_2000"] = {
        "bow_transform_k10_L6_ms_gpu": med(lambda: V.transform(da, 4), a.reps),
        "bow_transform_k10_L6_ms_cpu_oracle_1thread": med(lambda: OV.transform(da, 4), 5),
        "fuse_6000pts_ms_gpu": med(lambda: orbfe.Fuse(FA, *fu), a.reps),
        "fuse_6000pts_ms_cpu_oracle_1thread": med(lambda: O.fuse(OFA, *fu), 5),
        "search_by_projection_sim3_6000pts_ms_gpu": med(lambda: orbfe.SearchByProjectionSim3(FA, *sp), a.reps),
        "search_by_projection_sim3_6000pts_ms_cpu_oracle_1thread": med(lambda: O.search_by_projection_sim3(OFA, *sp), 5),
        "search_by_bow_keyframes_ms_gpu": med(lambda: orbfe.SearchByBoWKeyFrames(FB, *bk), a.reps),
        "search_by_bow_keyframes_ms_cpu_oracle_1thread": med(lambda: O.search_by_bow_keyframes(OFB, *bk), 5),
        "search_for_triangulation_ms_gpu": med(lambda: orbfe.SearchForTriangulation(FB, *tr), a.reps),
        "search_for_triangulation_ms_cpu_oracle_1thread": med(lambda: O.search_for_triangulation(OFB, *tr), 5),
        "note": "wall clock of one C-ABI call through the ctypes wrapper, host arrays in/out (H2D/D2H included; the feature "
                "vectors are passed flat, as the C++ shim passes them)"}
    # the per-frame tracking front-end end to end, device-resident: stereo pair in -> extraction -> stereo matching -> matcher view
    # (device to device) -> SearchLocalPoints over 5000 local map points; only the images and the map points cross PCIe
    lt, rt = synth.stereo_pair(seed=71)
    ext = orbfe.ORBextractor(lib=L, max_images=2)
    rngt = np.random.default_rng(72)
    def extract_pair():
        ext.upload([lt, rt]); ext.run(2); ext.run_stereo(1, P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"])
    extract_pair()
    bt = ext.download(2, ext.make_buffers(2, stereo=True))
    nt = int(bt["n"][0]); kt, dt = bt["kps"][0, :nt].copy(), bt["desc"][0, :nt].copy()
    fxk, cxk, cyk = np.float32(718.856), np.float32(607.1928), np.float32(185.2157)
    src = rngt.integers(0, nt, 5000)
    zt = rngt.uniform(4, 60, 5000).astype(np.float32)
    wt = np.stack([(kt["x"][src] - cxk) / fxk * zt, (kt["y"][src] - cyk) / fxk * zt, zt], 1).astype(np.float32)
    dist = np.linalg.norm(wt, axis=1).astype(np.float32)
    nrm = (wt / dist[:, None]).astype(np.float32)
    sc = ext.GetScaleFactors()
    raw = (dist * sc[kt["octave"][src]]).astype(np.float32)
    mx, mn = (np.float32(1.2) * raw).astype(np.float32), (np.float32(0.8) * raw / sc[-1]).astype(np.float32)
    mdesc = dt[src].copy(); hob = np.ones(5000, np.uint8); occ = np.zeros(nt, np.uint8)
    lsf = float(np.log(np.float32(1.2)).astype(np.float32))
    I3, Z3 = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    Fd = orbfe.Frame.from_extractor(ext, (0.0, 1241.0, 0.0, 376.0), slot=0, stereo=True)
    def track_frame():
        extract_pair()
        Fd.refresh_from_extractor(ext, slot=0, stereo=True)   # steady state: the handle's device arrays are reused
        return orbfe.SearchLocalPoints(Fd, wt, nrm, mn, mx, raw, I3, Z3, Z3, float(fxk), float(fxk), float(cxk), float(cyk), P.KITTI["bf"], lsf,
                                       mdesc, hob, occ, 1, 0.8)
    nv, nm = track_frame()[:2]
    oL_, oR_ = O.Extractor(), O.Extractor()
    def track_frame_cpu():
        okl, odl = oL_.extract(lt); okr, odr = oR_.extract(rt)
        _, our, _ = O.stereo_match(oL_, oR_, okl, odl, okr, odr, P.KITTI["bf"], P.KITTI["bf"] / P.KITTI["fx"])
        OFt = O.Frame(okl, odl, sc, (0.0, 1241.0, 0.0, 376.0), our)
        _, tr = O.is_in_frustum(wt, nrm, mn, mx, raw, I3, Z3, Z3, float(fxk), float(fxk), float(cxk), float(cyk), P.KITTI["bf"],
                                (0.0, 1241.0, 0.0, 376.0), lsf, 8, 0.5)
        return O.search_by_projection_mappoints(OFt, tr["in_view"], tr["proj_x"], tr["proj_y"], tr["proj_xr"], tr["level"], tr["view_cos"],
                                                mdesc, hob, occ, 1, 0.8)
    onm = track_frame_cpu()[0]
    assert nm == onm, (nm, onm)
    out["tracking_frontend_device_resident"] = {
        "stereo_pair_to_local_map_matches_ms_gpu": med(track_frame, a.reps),
        "same_on_cpu_oracle_1thread_ms": med(track_frame_cpu, 3),
        "local_map_points": 5000, "in_view": nv, "matches": nm,
        "path": "orbfe_upload(2) + orbfe_run + orbfe_run_stereo + orbfe_frame_refresh_from_extractor + orbfe_search_local_points"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
