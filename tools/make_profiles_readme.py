#!/usr/bin/env python
"""Regenerates profiles/README.md from the committed round artefacts (bench JSON lines, ncu launch list,
ncu --set full summary).  usage: python tools/make_profiles_readme.py r01"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
P = os.path.join(ROOT, "profiles")
def jline(name):
    with open(os.path.join(P, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])
d = jline(f"{tag}_bench_n1.json")
out = [f"# profiles/ — {tag} evidence\n",
       "All numbers: one B200 (gpurun box, SM clock 1965 MHz, no throttle reasons), `python bench.py --steps 30 --warmup 3`:",
       "64 synthetic KITTI-shaped stereo pairs (128 frames, 1241x376) per step, nFeatures 2000, 8 levels, FAST 20/7.\n",
       "| quantity | value |", "|---|---|",
       f"| `value` (inputs resident, CUDA events) | **{d['value']:.0f} stereo pairs/s** ({d['ms_per_step']:.3f} ms / 64-pair step) |",
       f"| `e2e` (pinned host in -> host out, 2 handles pipelined) | **{d['e2e']['value']:.0f} stereo pairs/s** ({d['e2e']['h2d_bytes_per_step']/1e6:.1f} MB H2D + {d['e2e']['d2h_bytes_per_step']/1e6:.1f} MB D2H per step) |",
       f"| p50 latency, one pair, drop-in calls driven from Python/ctypes (2x `orbfe_extract` on 2 threads + `orbfe_stereo_match`) | {d['latency']['p50_ms_per_frame']:.3f} ms |",
       f"| p50 latency, one pair, drop-in path driven from C++ (`orbfe_shim.hpp`: 2 `std::thread`s x `ORBextractor::Compute` + `ComputeStereoMatches`) | {d['latency'].get('cpp_shim_p50_ms_per_frame', float('nan')):.3f} ms |",
       f"| p50 latency, one pair, one batched call sequence | {d['latency'].get('fused_p50_ms_per_frame', float('nan')):.3f} ms |",
       f"| CPU baseline (oracle port, {d['cpu_baseline']['cores']} host cores) | {d['cpu_baseline']['value']:.0f} stereo pairs/s ({d['cpu_baseline']['sample']}) |",
       f"| kernel launches in the timed region | {d['gpu_launches']} ({d['gpu_launches']//d['steps']} per step) |",
       f"| keypoints / image, stereo matches / pair | {d['keypoints_per_image']:.0f}, {d['stereo_matches_per_pair']:.0f} |", ""]
out += ["## Per-stage CUDA-event times inside the timed region (roofline.stages)\n",
        "| stage | ms / step | share | algorithmic MB / step | GB/s | fraction of measured HBM peak (6523 GB/s) |", "|---|---|---|---|---|---|"]
for k, v in d["roofline"]["stages"].items():
    ab = v.get("alg_bytes_per_step")
    if ab:
        out.append(f"| {k} | {v['ms_per_step']:.3f} | {100*v['share']:.1f} pct | {ab/1e6:.0f} | {v['gbs']:.0f} | {v['frac_of_hbm_peak']:.3f} |")
    else:
        out.append(f"| {k} | {v['ms_per_step']:.3f} | {100*v['share']:.1f} pct | - | - | - |")
ws = d["roofline"]["whole_step"]
out += ["", f"Whole step: {ws['alg_bytes']/1e6:.0f} MB algorithmic (I + 4P per image) -> {ws['gbs']:.0f} GB/s = {ws['frac']:.3f} of the measured HBM peak.",
        f"Dominant kernel: `{d['roofline']['kernel']}` ({d['roofline']['launch_ms']:.3f} ms / launch, {d['roofline']['alg_bytes_per_launch']/1e6:.0f} MB algorithmic, "
        f"ncu DRAM traffic {d['roofline']['traffic']/1e6 if d['roofline'].get('traffic') else float('nan'):.0f} MB / launch): it is ALU-bound (see the ncu summary), not HBM-bound.", ""]
sc = []
for n in (1, 2, 4, 8):
    fn = os.path.join(P, f"{tag}_bench_n{n}.json")
    if os.path.exists(fn):
        sc.append((n, jline(f"{tag}_bench_n{n}.json")))
if len(sc) > 1:
    v1, e1 = sc[0][1]["value"], sc[0][1]["e2e"]["value"]
    out += ["## Weak scaling on one box (torchrun, one rank per GPU, frames sharded, no collective)\n",
            "| GPUs | `value` pairs/s | x N=1 | `e2e` pairs/s | x N=1 |", "|---|---|---|---|---|"]
    for n, dd in sc:
        out.append(f"| {n} | {dd['value']:.0f} | {dd['value']/v1:.2f} | {dd['e2e']['value']:.0f} | {dd['e2e']['value']/e1:.2f} |")
    out += ["", "The kernels scale linearly (ranks share nothing).  The end-to-end leg moves 77 MB per 64-pair step per GPU through the host: on this",
            "box (a 32-vCPU KVM guest, all 8 GPUs behind one NUMA node) it saturates at about 90-130 k pairs/s = 110-160 GB/s of combined",
            "H2D + D2H traffic, whatever the number of GPUs; that is the host's limit, not the GPUs'.", ""]
ll = os.path.join(P, f"{tag}_launches_pairs16.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]; ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
    tot = sum(float(r[vi]) for r in rows[h + 1:])
    out += ["## ncu launch list of one step (`--pairs 16`, cold-cache, serialised: compare SHARES)\n", "| kernel | grid | us | share |", "|---|---|---|---|"]
    for r in rows[h + 1:]:
        out.append(f"| `{r[ki].split('(')[0]}` | {r[gi]} | {float(r[vi])/1e3:.1f} | {100*float(r[vi])/tot:.1f} % |")
    out.append("")
out += ["## Files\n", "| file | what |", "|---|---|"]
desc = {"_bench_n1.json": "bench.py JSON line, N=1", "_bench_n2.json": "bench.py JSON line, N=2 (torchrun, weak scaling)",
        "_bench_n4.json": "bench.py JSON line, N=4", "_bench_n8.json": "bench.py JSON line, N=8",
        "_bench_reference.json": "`bench.py --impl reference` JSON line (oracle port on the host cores)",
        "_launches_pairs16.csv": "`ncu --metrics gpu__time_duration.sum --clock-control none` launch list of `bench.py --steps 2 --warmup 3 --pairs 16 --no-cpu --no-latency`",
        "_matchers.json": "`tests/bench_matchers.py`: matcher rows (M2-M4), N1 routines and the N3 vocabulary transform, wall clock of one C-ABI call on the GPU next to the CPU oracle (1 thread)",
        "_ncu_full_summary.txt": "`ncu --set full --clock-control none --import-source on` of the same command, one step, summarised by tools/ncu_summary.py (time, DRAM bytes, occupancy, pipe utilisation, stall reasons per kernel)",
        "_ncu_sass_regions_fast.txt": "hot SASS regions of `k_fast_cells` (tools/ncu_sass_regions.py)"}
for f in sorted(os.listdir(P)):
    if f == "README.md":
        continue
    what = next((v for k, v in desc.items() if f.endswith(k)), "earlier evidence kept for the record (first kernel versions)")
    out.append(f"| `{f}` | {what} |")
open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:20]))
