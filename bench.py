#!/usr/bin/env python
"""bench.py -- stereo frames/s of the ORB front-end hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over the offline sequence of BASELINE.json configs[2]: 4541 synthetic KITTI-shaped
stereo pairs (1241x376 u8), sharded frame-wise across the ranks (slam_framework_b200/shard.py) and processed in batches of
`--pairs` pairs: ORB extraction of both images (8-level pyramid, grid FAST, quad-tree, blur, orientation + rBRIEF) +
Frame::ComputeStereoMatches, on one GPU per rank.  Frames are independent, so ranks share nothing (no collective on the data
path); the job is fixed, so adding GPUs is strong scaling.

  value : stereo pairs/s with the batch already resident in HBM (device-side CUDA events on the
          library's own stream, max over ranks).
  e2e   : the same metric through the C-ABI calls a user makes (orbfe_upload from pinned host
          memory -> orbfe_run -> orbfe_run_stereo -> orbfe_download to host arrays), host<->device
          copies inside the timed region, wall clock, max over ranks.
  roofline      : per-stage CUDA-event times of the timed region against the measured HBM peak.
  cpu_baseline  : the oracle port (CPU restatement of the reference path) on the box's host cores, bounded sample; rank 0 at
                  N=1 only.  The reference's own sources built against the OpenCV stand-in (oracle/_ref) are timed beside it
                  (`reference_build_value`); the faster of the two is the baseline.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H = 1241, 376
NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH = 2000, 1.2, 8, 20, 7   # config/kitti_config_stereo.json
BF, FX = 386.1448, 718.856
METRIC = "stereo frames/s, ORB extract+stereo match 1241x376"
UNIT = "stereo pairs/s"


def level_pixels(w, h, nlevels=NLEVELS):
    """sum of pyramid level pixels, sizes exactly as orb_extractor.cpp:1055-1056"""
    s, tot = np.float32(1.0), 0
    for l in range(nlevels):
        inv = np.float32(1.0) / s
        lw = int(np.rint(np.float32(w) * inv)); lh = int(np.rint(np.float32(h) * inv))
        tot += lw * lh
        s = np.float32(np.float64(s) * np.float64(np.float32(SCALE)))
    return tot


def algorithmic_bytes_per_image(w=W, h=H):
    """compulsory traffic per image, u8 = 1 B/px."""
    I, P = w * h, level_pixels(w, h)
    return {"pyramid": I + P, "fast": P, "blur": 2 * P, "extract_total": I + 4 * P}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + ",".join(self.FIELDS),
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(n_seq, pairs_per_batch, distinct, world):
    """`config` of the JSON line: the workload, identical for the GPU arm and the reference (CPU) arm of the same invocation"""
    n_img = 2 * pairs_per_batch
    lo, hi = 0, -(-n_seq // world)
    return {"workload": f"configs[2]: offline batch of {n_seq} synthetic KITTI-size stereo pairs (1241x376; configs[1] per pair: "
                        "L+R ORB extraction, nFeatures 2000, 8 levels, 1.2, FAST 20/7, + ComputeStereoMatches), sharded "
                        "frame-wise across the ranks, processed in batches",
            "step": "one pass over the whole sequence (each rank: its contiguous shard)", "sequence_pairs": n_seq,
            "pairs_per_batch": pairs_per_batch, "batches_per_step_rank0": -(-min(hi, n_seq) // pairs_per_batch), "distinct_pairs": distinct,
            "parallelism": f"frames sharded x{world}, no collective",
            "l2": "per-batch working set (images+pyramids+blur) ~%d MB > 126 MB L2" % ((n_img * (H * W + 2 * 1738559)) >> 20)}


def make_pairs(n, seed0):
    from slam_framework_b200 import synth
    return [synth.stereo_pair(H, W, seed=seed0 + i) for i in range(n)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_throughput(pairs, n_pairs, threads, pair_threads=1):
    """oracle batch driver (oracle/orb_oracle_batch.cpp) over n_pairs pairs cycling through `pairs`."""
    import oracle_lib as O
    L = O.lib()
    lp = (ctypes.c_void_p * n_pairs)()
    rp = (ctypes.c_void_p * n_pairs)()
    for i in range(n_pairs):
        l, r = pairs[i % len(pairs)]
        lp[i], rp[i] = l.ctypes.data, r.ctypes.data
    kps, mt = ctypes.c_long(), ctypes.c_long()
    sec = L.orc_bench_stereo_batch(lp, rp, n_pairs, W, H, NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, BF, BF / FX,
                                   threads, pair_threads, ctypes.byref(kps), ctypes.byref(mt))
    return n_pairs / sec, sec, kps.value, mt.value


def ref_build_throughput(pairs, n_pairs, workers):
    """the reference's OWN sources (oracle/_ref/libslam_ref.so: Frame's stereo constructor = 2 extraction threads +
    ComputeStereoMatches per pair) over n_pairs pairs; None when that library is not present"""
    try:
        import reference_lib as R
        if not R.available():
            return None
        L = R.lib()
    except Exception:
        return None
    L.ref_bench_stereo_batch.restype = ctypes.c_double
    L.ref_bench_stereo_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_float] + [ctypes.c_int] * 3 + \
        [ctypes.c_float] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lp = (ctypes.c_void_p * n_pairs)()
    rp = (ctypes.c_void_p * n_pairs)()
    for i in range(n_pairs):
        l, r = pairs[i % len(pairs)]
        lp[i], rp[i] = l.ctypes.data, r.ctypes.data
    kps, mt = ctypes.c_long(), ctypes.c_long()
    sec = L.ref_bench_stereo_batch(lp, rp, n_pairs, W, H, NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, FX, FX, 607.1928, 185.2157, BF,
                                   workers, ctypes.byref(kps), ctypes.byref(mt))
    return n_pairs / sec


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference's own
    sources do compile here against a stand-in OpenCV layer (oracle/_ref), but that build runs our
    scalar stand-ins for OpenCV's SIMD primitives and is ~1.5x slower than the oracle port; the port (bit-identical to it)
    is timed, with one worker per host core over independent pairs: the conservative baseline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    pairs = make_pairs(min(8, max(2, cores)), 0)
    per_step = 8 * max(cores, 4)  # ~0.4 s of host work per step: thread start-up is amortised
    for _ in range(max(args.warmup, 1)):
        cpu_reference_throughput(pairs, per_step, cores)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cpu_reference_throughput(pairs, per_step, cores)
        done += per_step
    sec = time.perf_counter() - t0
    val = done / sec
    sample = f"{per_step} stereo pairs per step x {args.steps} steps, one oracle worker per host core"
    # the reference's own sources (oracle/_ref), timed beside the port on a smaller sample: workers x 2 extraction threads
    ref_build = ref_build_throughput(pairs, 2 * max(cores, 4), max(cores // 2, 1))
    kind = "port"
    if ref_build is not None and ref_build > val:  # report the faster of the two CPU implementations
        val, kind = ref_build, "reference"
        sample = f"{2 * max(cores, 4)} stereo pairs through oracle/_ref (the reference's own sources), {max(cores // 2, 1)} workers x 2 threads"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(args.sequence, args.pairs, min(args.pairs, args.distinct), max(args.gpus, 1)),
           "reference_sample": {"pairs_per_step": per_step, "note": "each step of this arm is a bounded sample of the sequence "
                                "(the same synthetic pairs, seeds 0..) on the host cores"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                            "port_value": done / sec, "reference_build_value": ref_build,
                            "note": "port = oracle restatement; reference_build = the reference's own sources compiled against the "
                                    "OpenCV stand-in (oracle/_ref); the two are bit-identical, the faster one is the baseline"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def p50_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(t)


def config_figures(L, local, pairs, baseline):
    """BASELINE.json configs 1, 2, 4, 5: p50 wall time of the drop-in calls (host arrays in, host arrays out) on the GPU next
    to the same work on the CPU (oracle port; the reference's threading: 2 threads for a stereo pair, frame.cpp:86-89, one
    otherwise), each GPU figure with its own clocks sample.  Rank 0, N = 1 only."""
    import oracle_lib as O
    import parity_common as P
    from slam_framework_b200 import orbfe, synth
    out = {}

    def clocked(fn):
        smp = ClockSampler(local)
        smp.start()
        try:
            v = fn()
        finally:
            c = smp.stop()
        return v, c

    l, r = pairs[0]
    # ---- config 1: one 1241x376 frame, nFeatures 2000
    e1 = orbfe.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, lib=L)
    o1 = O.Extractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH)
    g, c = clocked(lambda: p50_ms(lambda: e1.Compute(l), 60, 10))
    out["config1_single_frame"] = {"gpu_p50_ms": g, "cpu_p50_ms": p50_ms(lambda: o1.extract(l), 7, 1), "cpu_threads": 1,
                                   "path": "ORBextractor::Compute = orbfe_extract (host image in, keypoints + descriptors out)", "clocks": c}
    # ---- config 2: one stereo pair = 2 x Compute on two host threads + ComputeStereoMatches
    e2 = orbfe.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, lib=L)

    def pair_gpu():
        res = [None, None]
        th = threading.Thread(target=lambda: res.__setitem__(1, e2.Compute(r)))
        th.start()
        res[0] = e1.Compute(l)
        th.join()
        orbfe.ComputeStereoMatches(e1, e2, res[0][0], res[0][1], res[1][0], res[1][1], BF, baseline)
    g, c = clocked(lambda: p50_ms(pair_gpu, 60, 10))
    cfg2 = {"gpu_p50_ms": g, "path": "2 x orbfe_extract on two host threads + orbfe_stereo_match (python / ctypes driver)", "clocks": c}
    ef = orbfe.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_images=2, lib=L)
    bf2 = ef.make_buffers(2, stereo=True)

    def pair_fused():
        ef.upload([l, r]); ef.run(2); ef.run_stereo(1, BF, baseline); ef.download(2, bf2)
    cfg2["gpu_fused_p50_ms"] = p50_ms(pair_fused, 60, 10)
    cfg2["gpu_fused_path"] = "orbfe_upload(2) + orbfe_run + orbfe_run_stereo + orbfe_download on one handle"
    try:  # the drop-in path driven from C++ (include/orbfe_shim.hpp): what the reference's Frame constructor runs
        import tempfile
        from test_shim_cpp import build_demo
        exe = build_demo()
        with tempfile.TemporaryDirectory() as td:
            l.tofile(os.path.join(td, "l.raw")); r.tofile(os.path.join(td, "r.raw"))
            o = subprocess.check_output([exe, "--latency", str(W), str(H), os.path.join(td, "l.raw"), os.path.join(td, "r.raw"), "200"],
                                        text=True, timeout=120).split()
        cfg2["gpu_cpp_shim_p50_ms"] = float(o[0])
        cfg2["gpu_cpp_shim_path"] = "C++ shim: 2 std::threads x ORBextractor::Compute + orbfe::ComputeStereoMatches, 200 frames"
    except Exception as e:  # the demo binary is test infrastructure: its absence must not fail the bench
        cfg2["gpu_cpp_shim_error"] = str(e)[:200]
    cfg2["cpu_p50_ms"] = p50_ms(lambda: cpu_reference_throughput([pairs[0]], 1, 1, 2), 7, 1)
    cfg2["cpu_threads"] = 2
    cfg2["cpu_path"] = "oracle port, the pair's two extractions on two threads (frame.cpp:86-89), then the stereo match"
    try:
        import reference_lib as R
        if R.available():
            cfg2["cpu_reference_build_p50_ms"] = p50_ms(lambda: R.Frame(l, r), 5, 1)
            cfg2["cpu_reference_build_path"] = "the reference's own Frame stereo constructor (oracle/_ref), 2 extraction threads"
    except Exception:
        pass
    out["config2_stereo_pair"] = cfg2
    e2.close(); ef.close()
    # ---- config 4: monocular initialisation: 2 x extract(4000) + SearchForInitialization(window 100)
    a_img, b_img = synth.shifted_frame(21, dx=8, dy=4)
    e4 = orbfe.ORBextractor(2 * NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, lib=L)
    o4 = O.Extractor(2 * NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH)
    m4 = orbfe.OrbMatcher(0.9, True)
    sc = e4.GetScaleFactors()

    def init_gpu():
        k1, d1 = e4.Compute(a_img); k2, d2 = e4.Compute(b_img)
        F1 = orbfe.Frame(k1, d1, sc, (0, W, 0, H), lib=L); F2 = orbfe.Frame(k2, d2, sc, (0, W, 0, H), lib=L)
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        n = m4.SearchForInitialization(F1, F2, prev, 100)[0]
        F1.close(); F2.close()
        return n

    def init_cpu():
        k1, d1 = o4.extract(a_img); k2, d2 = o4.extract(b_img)
        F1 = O.Frame(k1, d1, sc, (0.0, float(W), 0.0, float(H))); F2 = O.Frame(k2, d2, sc, (0.0, float(W), 0.0, float(H)))
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        return O.search_for_initialization(F1, F2, prev, 100, 0.9, True)[0]
    assert init_gpu() == init_cpu()
    g, c = clocked(lambda: p50_ms(init_gpu, 30, 5))
    out["config4_mono_init_4000"] = {"gpu_p50_ms": g, "cpu_p50_ms": p50_ms(init_cpu, 3, 1), "cpu_threads": 1, "matches": init_gpu(),
                                     "path": "2 x ORBextractor::Compute(4000) + 2 x frame view + SearchForInitialization", "clocks": c}
    e4.close()
    # ---- config 5: 1080p and 4K, extract(8000) + SearchByProjection against 20 000 projected map points
    for (h, w) in ((1080, 1920), (2160, 3840)):
        img = synth.frame(h, w, seed=w)
        e5 = orbfe.ORBextractor(4 * NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, lib=L)
        o5 = O.Extractor(4 * NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH)
        kps, desc = e5.Compute(img)
        mp = P.synth_map_points(kps, desc, np.random.default_rng(3), 20000)
        margs = (mp["valid"], mp["px"], mp["py"], mp["pxr"], mp["lvl"], mp["view"], mp["desc"], mp["has_obs"], mp["occupied"])
        m5 = orbfe.OrbMatcher(0.8)
        sc5 = e5.GetScaleFactors()

        def track_gpu():
            k, d = e5.Compute(img)
            F = orbfe.Frame(k, d, sc5, (0, w, 0, h), lib=L)
            n = m5.SearchByProjectionMapPoints(F, *margs, 1)[0]
            F.close()
            return n

        def track_cpu():
            k, d = o5.extract(img)
            return O.search_by_projection_mappoints(O.Frame(k, d, sc5, (0.0, float(w), 0.0, float(h))), *margs, 1, 0.8)[0]
        assert track_gpu() == track_cpu()
        g, c = clocked(lambda: p50_ms(track_gpu, 20, 3))
        out[f"config5_{w}x{h}_8000"] = {"gpu_p50_ms": g, "cpu_p50_ms": p50_ms(track_cpu, 2, 0), "cpu_threads": 1, "keypoints": int(len(kps)),
                                        "matches": track_gpu(), "map_points": 20000,
                                        "path": "ORBextractor::Compute(8000) + frame view + SearchByProjection(F, 20k map points, th 1)", "clocks": c}
        e5.close()
    e1.close()
    return out


SEQ_PAIRS = 4541   # BASELINE.json configs[2]: KITTI-00 length


def run_ours(args):
    import torch
    import torch.distributed as dist
    from slam_framework_b200 import orbfe, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa, orig_aff = None, None
    try:
        orig_aff = os.sched_getaffinity(0)
    except Exception:
        pass
    try:  # NUMA: run (and first-touch the pinned staging buffers) on the cores closest to this rank's GPU
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa = len(cpus)
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = orbfe.load()
    B = args.pairs
    n_img = 2 * B
    n_seq = args.sequence
    lo, hi = shard.shard_range(n_seq, rank, world)       # this rank's contiguous shard of the sequence
    n_local = hi - lo
    batches = [min(B, n_local - s) for s in range(0, n_local, B)]   # pairs per batch; the last one may be partial
    ex = orbfe.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_images=n_img, max_size=(W, H), lib=L)
    # synthetic input: `distinct` different pairs (the same on every rank: pair i of the sequence is pair i % distinct), staged in
    # pinned host memory.  With distinct == B every batch uploads the same B pairs: frames are independent, the copies are real.
    distinct = min(B, args.distinct)
    pairs = make_pairs(distinct, 0)
    host = torch.empty((n_img, H, W), dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    for p in range(B):
        l, r = pairs[p % distinct]
        hnp[2 * p], hnp[2 * p + 1] = l, r
    ptrs = (ctypes.c_void_p * n_img)(*[hnp[i].ctypes.data for i in range(n_img)])
    baseline = BF / FX

    def pinned_buffers(e):
        """make_buffers layout in pinned host memory (D2H lands here without a staging copy)"""
        cap = e.max_keypoints()
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        return dict(kps=mk((n_img, cap, 28), torch.uint8).view(orbfe.KP_DTYPE).reshape(n_img, cap),
                    desc=mk((n_img, cap, 32), torch.uint8), n=mk((n_img,), torch.int32), cap=cap,
                    ur=mk((n_img, cap), torch.float32), depth=mk((n_img, cap), torch.float32))

    def step_resident():
        for nb in batches:
            ex.run(2 * nb)
            ex.run_stereo(nb, BF, baseline)

    # ---- device-resident leg -----------------------------------------------------------------
    ex.upload_ptrs(ptrs, n_img, W, H, W)
    ex.run(n_img); ex.run_stereo(B, BF, baseline)   # sizes the pair table for the largest batch
    for _ in range(max(args.warmup, 3)):
        step_resident()
    ex.sync()
    ex.set_stage_timing(True)
    ex.stage_summary()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ex.launch_count()
    ex.event_record(0)
    for _ in range(args.steps):
        step_resident()
    ex.event_record(1)
    ex.sync()
    barrier()
    ms_local = ex.event_elapsed_ms(0, 1) if n_local else 0.0
    ms = max_over_ranks(ms_local)
    launches = ex.launch_count() - launches0
    stages, runs = ex.stage_summary()
    ex.set_stage_timing(False)
    value = n_seq * args.steps / (ms * 1e-3)

    # ---- end-to-end leg (host buffers, copies inside the timed region) ---------------------------
    # three handles (= three streams) take the batches in turn so that the H2D copy of batch k+1 and the D2H copy of batch
    # k-1 overlap the kernels of batch k; every batch uploads its own inputs and downloads its own results.
    NL = 3
    lanes = [(ex, pinned_buffers(ex))]
    for _ in range(NL - 1):
        e = orbfe.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_images=n_img, max_size=(W, H), lib=L)
        lanes.append((e, pinned_buffers(e)))

    def step_e2e(k0):
        for j, nb in enumerate(batches):
            e, b = lanes[(k0 + j) % NL]
            e.sync()  # the batch this handle took NL batches ago has delivered its results
            e.upload_ptrs(ptrs, 2 * nb, W, H, W)
            e.run(2 * nb)
            e.run_stereo(nb, BF, baseline)
            e.download_async(2 * nb, b)
        return k0 + len(batches)

    # the same pass with the inputs resident, batches taken in turn by the NL handles (no copies): what overlapping the
    # kernels of consecutive batches on several streams is worth on its own (extra key, not `value`: the per-stage CUDA-event
    # times above are only clean on a single stream)
    def step_resident_lanes(k0):
        for j, nb in enumerate(batches):
            e = lanes[(k0 + j) % NL][0]
            e.run(2 * nb)
            e.run_stereo(nb, BF, baseline)
        return k0 + len(batches)
    for e, _ in lanes[1:]:
        e.upload_ptrs(ptrs, n_img, W, H, W)
        e.run(n_img); e.run_stereo(B, BF, baseline)
    kk = step_resident_lanes(0)
    for e, _ in lanes:
        e.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        kk = step_resident_lanes(kk)
    for e, _ in lanes:
        e.sync()
    t_lanes = max_over_ranks(time.perf_counter() - t0)
    barrier()

    k = 0
    for _ in range(max(args.warmup, 3)):
        k = step_e2e(k)
    for e, _ in lanes:
        e.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        k = step_e2e(k)
    for e, _ in lanes:
        e.sync()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    # ---- copy-only ceiling of this box for the e2e leg: per batch ONE pinned H2D copy of the frames and ONE D2H copy of the
    # results, the two directions on separate streams, no kernels, all ranks at once (tools/copy_ceiling.py is the
    # stand-alone form).  e2e / ceiling says how much of the host <-> device path the pipeline uses.
    ceil_pairs = None
    try:
        cap0 = ex.max_keypoints()
        hb, db = n_img * H * W, n_img * cap0 * (28 + 32 + 4 + 4) + n_img * 4
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        h_in = [torch.empty(hb, dtype=torch.uint8, pin_memory=True) for _ in range(NL)]
        d_in = [torch.empty(hb, dtype=torch.uint8, device="cuda") for _ in range(NL)]
        h_out = [torch.empty(db, dtype=torch.uint8, pin_memory=True) for _ in range(NL)]
        d_out = [torch.empty(db, dtype=torch.uint8, device="cuda") for _ in range(NL)]

        def pump(k):
            for j in range(k):
                with torch.cuda.stream(s_in):
                    d_in[j % NL].copy_(h_in[j % NL], non_blocking=True)
                with torch.cuda.stream(s_out):
                    h_out[j % NL].copy_(d_out[j % NL], non_blocking=True)
        pump(10)
        barrier()
        iters = 150
        t0 = time.perf_counter()
        pump(iters)
        torch.cuda.synchronize()
        t_copy = max_over_ranks(time.perf_counter() - t0)
        barrier()
        ceil_pairs = world * B * iters / t_copy
        ceil_gbs = world * (hb + db) * iters / t_copy / 1e9
        del h_in, d_in, h_out, d_out
    except Exception:
        ceil_pairs = None
    buf = lanes[0][1]
    for e, _ in lanes[1:]:
        e.close()
    e2e_value = n_seq * args.steps / t_e2e
    cap = buf["cap"]
    n_batches = len(batches)
    h2d = 2 * n_local * H * W                                             # this rank, per step (= per pass over its shard)
    d2h = 2 * n_local * cap * (28 + 32 + 4 + 4) + 2 * n_local * 4 + 4 * n_batches
    nb_last = batches[-1] if batches else 0
    n_kp = int(buf["n"].sum())
    n_matched = int((buf["ur"][0::2] >= 0).sum())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (per-stage CUDA events on the library's stream inside the timed region, per 64-pair batch) ------
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes_per_image()
    kp_img = n_kp / n_img
    imgs_per_launch = 2.0 * n_local / max(n_batches, 1)   # average over the (possibly partial) batches of a pass
    stage_bytes = {  # algorithmic bytes per image for each stage (DESIGN.md "Kernels")
        "pyramid": alg["pyramid"], "fast": alg["fast"], "blur": alg["blur"],
        "quadtree": None, "describe": None, "stereo_search": None, "stereo_median": None}
    stage_launches = {"pyramid": NLEVELS, "fast": 1, "quadtree": 1, "blur": 1, "describe": 1, "stereo_search": 1,
                      "stereo_median": 1}
    per_stage = {}
    for kk, tot in stages.items():
        t_b = tot / max(runs, 1)   # per batch
        ent = {"ms_per_batch": t_b, "share": tot / max(sum(stages.values()), 1e-9), "launches_per_batch": stage_launches[kk]}
        if stage_bytes[kk] is not None and t_b > 0:
            ent["alg_bytes_per_batch"] = stage_bytes[kk] * imgs_per_launch
            ent["gbs"] = stage_bytes[kk] * imgs_per_launch / (t_b * 1e-3) / 1e9
            ent["frac_of_hbm_peak"] = ent["gbs"] / peak
        per_stage[kk] = ent
    dom = max(stages, key=lambda q: stages[q])
    dom_ms = stages[dom] / max(runs, 1)
    if stage_bytes[dom] is not None:
        dom_bytes = stage_bytes[dom] * imgs_per_launch
    else:  # integer/latency-bound stages: their compulsory HBM traffic is the candidate / keypoint records
        dom_bytes = {"quadtree": 4 * 13000 + 4 * kp_img, "describe": (4 + 28 + 32) * kp_img,
                     "stereo_search": 2 * (28 + 32) * kp_img, "stereo_median": 12 * kp_img}[dom] * imgs_per_launch
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    whole = alg["extract_total"] * 2 * n_local / (ms_local / args.steps * 1e-3) / 1e9 if ms_local > 0 else 0.0
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture; only when it was taken at this batch size
        import glob
        tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]
        tj = json.load(open(tf))
        traffic_src = f"{os.path.basename(tf)}: captured at {tj['n_images_in_capture']} images per launch, this run launches {n_img}"
        if dom in tj["dram_bytes_per_image"] and int(tj["n_images_in_capture"]) == n_img:
            traffic = tj["dram_bytes_per_image"][dom] * imgs_per_launch / stage_launches[dom]
        else:
            traffic_src += " -- different batch size or kernel, traffic not reported"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "the dominant kernel (grid FAST) is bound by instruction issue (integer pipes + shared-memory gathers), not by HBM: see DESIGN.md section 4 and profiles/",
                "alg_bytes_per_launch": dom_bytes / stage_launches[dom], "launch_ms": dom_ms / stage_launches[dom],
                "whole_step": {"alg_bytes": alg["extract_total"] * 2 * n_local, "gbs": whole, "frac": whole / peak},
                "stages": per_stage}

    # ---- BASELINE configs 1, 2, 4, 5: GPU p50 next to the CPU p50 ------------------------------------------------------
    configs = None
    if world == 1 and not args.no_latency:
        try:
            configs = config_figures(L, local, pairs, baseline)
        except Exception as e:  # secondary figures: a failure here must not lose the headline line
            configs = {"error": repr(e)[:300]}

    # ---- CPU baseline (oracle port on the host cores; bounded sample) ----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        if orig_aff:
            os.sched_setaffinity(0, orig_aff)  # the CPU baseline uses every host core again
        cores = host_cores()
        v1, s1, _, _ = cpu_reference_throughput(pairs, cores, cores)
        reps = int(min(max(12.0 / max(s1, 1e-3), 1), 400))  # ~12 s of host work
        n = cores * reps
        v, sec, ckps, cmt = cpu_reference_throughput(pairs, n, cores)
        rb = ref_build_throughput(pairs, 2 * max(cores, 4), max(cores // 2, 1))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} stereo pairs of the sequence ({sec:.1f} s), one oracle worker per host core", "port_value": v,
               "reference_build_value": rb,
               "note": "reference_build = the reference's own sources compiled against the OpenCV stand-in (oracle/_ref), "
                       "bit-identical to the port; the faster of the two is the baseline"}
        if rb is not None and rb > v:
            cpu["value"], cpu["kind"] = rb, "reference"

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic",
           "config": workload_config(n_seq, B, distinct, world), "host_affinity_cores": numa,
           "gpu_launches": int(launches),
           "value_multistream": {"value": n_seq * args.steps / t_lanes, "unit": UNIT, "streams": NL,
                                 "note": "inputs resident, batches issued round-robin on the handles the e2e leg uses (wall clock); "
                                         "`value` is the single-stream figure its per-stage times belong to"},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * t_e2e / args.steps, "lanes": NL,
                   "copy_ceiling": None if not ceil_pairs else {
                       "pairs_per_s": ceil_pairs, "combined_gbs": ceil_gbs, "fraction_of_ceiling": e2e_value / ceil_pairs,
                       "what": "the same per-batch H2D + D2H copies alone (pinned, full duplex, all ranks at once, no kernels), "
                               "measured in this run right after the e2e leg"},
                   "note": "bytes are rank 0's per step (= per pass over its shard); pinned host buffers both ways"},
           "clocks": clocks, "roofline": roofline,
           "keypoints_per_image": kp_img, "stereo_matches_per_pair": n_matched / max(nb_last if n_batches == 1 else B, 1)}
    if cpu:
        out["cpu_baseline"] = cpu
    if configs:
        out["configs"] = configs
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="stereo pairs per step per GPU")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic pairs tiled over the sequence")
    ap.add_argument("--sequence", type=int, default=SEQ_PAIRS, help="stereo pairs in the offline sequence (whole job, all ranks)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
