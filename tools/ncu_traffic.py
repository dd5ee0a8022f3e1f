#!/usr/bin/env python
"""DRAM traffic per stage from an `ncu --set full` report of tools/profile_step.py (one batch):
dram__bytes_read.sum + dram__bytes_write.sum of every launch, summed per stage, divided by the images in the batch.
usage: ncu_traffic.py report.ncu-rep n_images out.json"""
import csv, json, subprocess, sys
rep, n_img, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
ki, ri, wi, ii = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("smsp__inst_executed.sum")
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = [("k_pyramid", "pyramid"), ("k_fast", "fast"), ("k_octree", "quadtree"), ("k_blur", "blur"), ("k_orient", "describe"),
            ("k_stereo_rows", "stereo_search"), ("k_stereo_search", "stereo_search"), ("k_stereo_median", "stereo_median")]
tr, ins = {}, {}
for r in rows[2:]:
    st = next((s for k, s in stage_of if k in r[ki]), None)
    if st is None:
        continue
    b = float(r[ri]) * mult[U[ri]] + float(r[wi]) * mult[U[wi]]
    tr[st] = tr.get(st, 0.0) + b
    ins[st] = ins.get(st, 0.0) + float(r[ii])
json.dump({"_source": f"{rep.split('/')[-1]}: ncu --set full of tools/profile_step.py, one batch of {n_img} images (the batch bench.py times); "
                      "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per stage",
           "n_images_in_capture": n_img,
           "dram_bytes_per_image": {k: v / n_img for k, v in tr.items()},
           "warp_instructions_per_image": {k: v / n_img for k, v in ins.items()}}, open(out, "w"), indent=1)
print(open(out).read())
