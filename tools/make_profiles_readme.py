#!/usr/bin/env python
"""Regenerates profiles/README.md from the committed round-2 artefacts (bench JSON lines, ncu launch lists, matcher timings,
copy ceiling).  usage: python tools/make_profiles_readme.py [bench_n1.json] [launches.csv]
Defaults: the newest profiles/r02*_bench_n1.json and profiles/r02*_launches_pairs64.csv."""
import csv, glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def jline(path):
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def newest(pattern):
    c = sorted(glob.glob(os.path.join(P, pattern)))
    return c[-1] if c else None


bench = sys.argv[1] if len(sys.argv) > 1 else newest("r02*_bench_n1.json")
launches = sys.argv[2] if len(sys.argv) > 2 else newest("r02*_launches_pairs64.csv")
d = jline(bench)
cc = d["e2e"].get("copy_ceiling")
out = ["# profiles/ — round 2 evidence\n",
       f"Headline line: `{os.path.basename(bench)}` = `python bench.py` on one B200 (gpurun box, SM clock {d['clocks']['sm_mhz']:.0f} MHz, throttle reasons "
       f"{d['clocks']['reasons'] or 'none'}).  A step = one pass over {d['config']['sequence_pairs']} synthetic KITTI-shaped stereo pairs "
       f"(1241x376, nFeatures 2000, 8 levels, FAST 20/7) in {d['config']['pairs_per_batch']}-pair batches; {d['steps']} steps timed.\n",
       "| quantity | value |", "|---|---|",
       f"| `value` (inputs resident, CUDA events) | **{d['value']:.0f} stereo pairs/s** ({d['ms_per_step']:.1f} ms per pass) |",
       f"| `e2e` (pinned host in -> host out, {d['e2e']['lanes']} handles taking the batches in turn) | **{d['e2e']['value']:.0f} stereo pairs/s** "
       f"({d['e2e']['h2d_bytes_per_step']/1e9:.2f} GB H2D + {d['e2e']['d2h_bytes_per_step']/1e9:.2f} GB D2H per pass) |"]
if cc:
    out.append(f"| the same copies alone (no kernels), measured in the same run | {cc['pairs_per_s']:.0f} pairs/s ({cc['combined_gbs']:.1f} GB/s): e2e is at {cc['fraction_of_ceiling']:.2f} of it |")
if "cpu_baseline" in d:
    c = d["cpu_baseline"]
    out.append(f"| CPU baseline ({c['kind']}, {c['cores']} host cores) | {c['value']:.0f} stereo pairs/s ({c['sample']}); the reference's own sources: {c.get('reference_build_value') or float('nan'):.0f} |")
out += [f"| kernel launches in the timed region | {d['gpu_launches']} ({d['gpu_launches'] / d['steps'] / d['config']['batches_per_step_rank0']:.0f} per batch) |",
        f"| keypoints / image, stereo matches / pair | {d['keypoints_per_image']:.0f}, {d['stereo_matches_per_pair']:.0f} |", ""]
out += ["## Per-stage CUDA-event times inside the timed region (`roofline.stages`, per 64-pair batch = 128 frames)\n",
        "| stage | ms / batch | share | algorithmic MB / batch | GB/s | fraction of the measured HBM peak (%.0f GB/s) |" % d["roofline"]["peak"], "|---|---|---|---|---|---|"]
for k, v in d["roofline"]["stages"].items():
    ab = v.get("alg_bytes_per_batch")
    if ab:
        out.append(f"| {k} | {v['ms_per_batch']:.3f} | {100*v['share']:.1f} % | {ab/1e6:.0f} | {v['gbs']:.0f} | {v['frac_of_hbm_peak']:.3f} |")
    else:
        out.append(f"| {k} | {v['ms_per_batch']:.3f} | {100*v['share']:.1f} % | - | - | - |")
r = d["roofline"]
ws = r["whole_step"]
out += ["", f"Whole pass: {ws['alg_bytes']/1e9:.1f} GB algorithmic (I + 4P per image) -> {ws['gbs']:.0f} GB/s = {ws['frac']:.3f} of the measured HBM peak.",
        f"Dominant kernel: `{r['kernel']}` ({r['launch_ms']:.3f} ms / launch, {r['alg_bytes_per_launch']/1e6:.0f} MB algorithmic, ncu DRAM traffic "
        f"{(r['traffic'] or float('nan'))/1e6:.0f} MB / launch [{r.get('traffic_source')}]): bound by instruction issue, not by HBM (DESIGN.md section 4).", ""]
cfg = d.get("configs")
if cfg and "error" not in cfg:
    out += ["## BASELINE.json configs 1, 2, 4, 5: p50 wall time of the drop-in calls, GPU next to CPU\n", "| config | GPU p50 ms | CPU p50 ms (threads) | path |", "|---|---|---|---|"]
    for k, v in cfg.items():
        extra = "".join(f"; {a.replace('gpu_', '').replace('_p50_ms', '')}: {v[a]:.3f}" for a in ("gpu_fused_p50_ms", "gpu_cpp_shim_p50_ms") if a in v)
        out.append(f"| {k} | {v['gpu_p50_ms']:.3f}{extra} | {v['cpu_p50_ms']:.1f} ({v['cpu_threads']}) | {v['path']} |")
    out.append("")
sc = [(n, jline(f)) for n in (1, 2, 4, 8) for f in [newest(f"r02g_bench_n{n}.json")] if f]
ceil = newest("r02*_copy_ceiling_8gpu.json")
if len(sc) > 1 and ceil:
    cj = {r_["gpus"]: r_ for r_ in json.load(open(ceil))["runs"]}
    sc8 = [(n, jline(os.path.join(P, f"r02g_bench_n{n}.json"))) for n in (1, 2, 4, 8) if os.path.exists(os.path.join(P, f"r02g_bench_n{n}.json"))]
    v1, e1 = sc8[0][1]["value"], sc8[0][1]["e2e"]["value"]
    out += ["## Strong scaling on one 8-GPU box (`r02g_bench_n*.json`: torchrun, one rank per GPU, the 4541-pair sequence sharded, no collective)\n",
            "| GPUs | `value` pairs/s | x N=1 | `e2e` pairs/s | x N=1 | copies alone (in-run) | copies alone (`tools/copy_ceiling.py`) |", "|---|---|---|---|---|---|---|"]
    for n, dd in sc8:
        c2 = dd["e2e"].get("copy_ceiling") or {}
        out.append(f"| {n} | {dd['value']:.0f} | {dd['value']/v1:.2f} | {dd['e2e']['value']:.0f} | {dd['e2e']['value']/e1:.2f} | {c2.get('pairs_per_s', float('nan')):.0f} | "
                   f"{cj[n]['pairs_per_s']:.0f} ({cj[n]['combined_gbs']:.0f} GB/s) |")
    out += ["", "The kernels scale linearly (ranks share nothing).  End to end each rank moves 59.7 MB in and 17.6 MB out per 64-pair batch; the copies ALONE",
            "saturate at about 90 k pairs/s on 2-4 GPUs and 125 k on 8 (150 GB/s through the host), and the pipeline runs at that ceiling (the ratio",
            "exceeds 1 by the run-to-run spread of the copy measurement): the limit at N >= 2 is the host-device path of the box, not the GPUs.", ""]
if launches:
    rows = list(csv.reader(open(launches)))
    h = [i for i, r_ in enumerate(rows) if r_ and r_[0] == "ID"][0]
    Hh = rows[h]; ki, vi, gi, bi = Hh.index("Kernel Name"), Hh.index("Metric Value"), Hh.index("Grid Size"), Hh.index("Block Size")
    tot = sum(float(r_[vi]) for r_ in rows[h + 1:])
    out += [f"## ncu launch list of one 64-pair batch (`{os.path.basename(launches)}`; cold-cache, serialised: compare SHARES)\n", "| kernel | grid | block | us | share |", "|---|---|---|---|---|"]
    for r_ in rows[h + 1:]:
        out.append(f"| `{r_[ki].split('(')[0]}` | {r_[gi]} | {r_[bi]} | {float(r_[vi])/1e3:.1f} | {100*float(r_[vi])/tot:.1f} % |")
    out.append("")
mt = newest("r02*_matchers.json")
if mt:
    m = json.load(open(mt))
    out += [f"## Matcher calls (`{os.path.basename(mt)}`: wall clock of one C-ABI call through the ctypes wrapper, host arrays in / out)\n", "| case | routine | GPU ms | CPU oracle ms (1 thread) |", "|---|---|---|---|"]
    for k, v in m.items():
        for kk, val in v.items():
            if kk.endswith("_ms_gpu"):
                cpu = v.get(kk.replace("_ms_gpu", "_ms_cpu_oracle_1thread"))
                out.append(f"| {k} | {kk[:-7]} | {val:.3f} | {'%.3f' % cpu if cpu else '-'} |")
    out.append("")
desc = {
    "_bench_n1.json": "bench.py JSON line, N=1", "_bench_n2.json": "bench.py JSON line, N=2 (torchrun)", "_bench_n4.json": "bench.py JSON line, N=4",
    "_bench_n8.json": "bench.py JSON line, N=8", "_bench_reference.json": "`bench.py --impl reference` JSON line (CPU arm on the box's host cores)",
    "_launches_pairs64.csv": "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` launch list of `tools/profile_step.py --pairs 64` (one batch, the size bench.py times)",
    "_launches_pairs1.csv": "the same for ONE stereo pair (the latency path)",
    "_matchers.json": "`tests/bench_matchers.py`: every matcher routine, the vocabulary transform and the device-resident tracking front-end, GPU next to the CPU oracle",
    "_ncu_full_summary_pairs64.txt": "`ncu --set full --clock-control none --import-source on` of `tools/profile_step.py --pairs 64`, summarised by tools/ncu_summary.py (time, DRAM bytes, occupancy, pipes, lane efficiency, bank conflicts, stall reasons per kernel)",
    "_ncu_full_summary_matchers.txt": "the same for the matcher / frame-tail / vocabulary kernels (`tests/bench_matchers.py --reps 1`)",
    "_ncu_full_summary_fast.txt": "`k_fast_cells` after the round-major queue", "_ncu_full_summary_describe.txt": "`k_orient_describe` with TMA-staged patches",
    "_ncu_sass_regions_fast.txt": "hot SASS regions of `k_fast_cells` (tools/ncu_sass_regions.py)",
    "_ncu_lines_fast.txt": "executed warp instructions of `k_fast_cells` per CUDA source line (tools/ncu_line_profile.py)",
    "_traffic.json": "DRAM bytes and warp instructions per image and stage from the 128-frame ncu capture (tools/ncu_traffic.py); bench.py reports `roofline.traffic` from it",
    "_sass_mnemonics.txt": "static SASS mnemonic counts per kernel (`UTMALDG` = TMA loads in FAST, blur, pyramid, describe)",
    "_copy_ceiling_8gpu.json": "`tools/copy_ceiling.py` on the 8-GPU box: the e2e leg's copies alone on 1, 2, 4, 8 GPUs", "_topology_8gpu.txt": "`nvidia-smi topo -m` of that box",
    "_run_sequence.json": "`tools/run_sequence.py`: the single-process multi-GPU driver"}
out += ["## Files\n", "| file | what |", "|---|---|"]
for f in sorted(os.listdir(P)):
    if f == "README.md":
        continue
    what = next((v for k, v in desc.items() if f.endswith(k)), None)
    if what is None:
        what = "round-1 evidence, kept for the record" if f.startswith("r01") else "see DESIGN.md"
    out.append(f"| `{f}` | {what} |")
open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
