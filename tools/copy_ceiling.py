#!/usr/bin/env python
"""Host<->device copy ceiling of the box for the end-to-end leg of bench.py: per iteration and per GPU ONE pinned
host->device copy of a 64-pair batch of frames (128 x 1241 x 376 B = 59.7 MB) and ONE device->host copy of its results
(17.6 MB: keypoints, descriptors, uR, depth at full capacity), on 1, 2, 4, 8 GPUs concurrently, no kernels.  The two
directions run on separate streams (full duplex), `--inflight` iterations deep.  Prints one JSON line:
pairs/s the copies alone would allow (64 pairs per iteration per GPU) and the GB/s behind it, per GPU count.

    python tools/copy_ceiling.py [--gpus 1,2,4,8] [--iters 200]

Single process, one host thread: the copies are asynchronous, the host only enqueues.  bench.py's e2e value divided by this
figure says how close the pipeline is to what the PCIe / host-memory path of the box can deliver."""
import argparse, json, time
import torch

H2D = 128 * 1241 * 376
D2H = 17_600_000


def run(devs, iters, inflight):
    st = []
    for d in devs:
        torch.cuda.set_device(d)
        st.append(dict(
            dev=d, s_in=torch.cuda.Stream(d), s_out=torch.cuda.Stream(d),
            h_in=[torch.empty(H2D, dtype=torch.uint8).pin_memory() for _ in range(inflight)],
            d_in=[torch.empty(H2D, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(inflight)],
            h_out=[torch.empty(D2H, dtype=torch.uint8).pin_memory() for _ in range(inflight)],
            d_out=[torch.empty(D2H, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(inflight)]))
    def pump(n):
        for k in range(n):
            for s in st:
                j = k % inflight
                with torch.cuda.stream(s["s_in"]):
                    s["d_in"][j].copy_(s["h_in"][j], non_blocking=True)
                with torch.cuda.stream(s["s_out"]):
                    s["h_out"][j].copy_(s["d_out"][j], non_blocking=True)
    def sync():
        for s in st:
            torch.cuda.synchronize(s["dev"])
    pump(10); sync()
    t0 = time.perf_counter()
    pump(iters); sync()
    dt = time.perf_counter() - t0
    n = len(devs)
    return {"gpus": n, "pairs_per_s": 64 * iters * n / dt, "h2d_gbs": H2D * iters * n / dt / 1e9, "d2h_gbs": D2H * iters * n / dt / 1e9,
            "combined_gbs": (H2D + D2H) * iters * n / dt / 1e9, "ms_per_iteration": 1e3 * dt / iters}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--inflight", type=int, default=3)
    a = ap.parse_args()
    have = torch.cuda.device_count()
    out = {"h2d_bytes_per_iteration_per_gpu": H2D, "d2h_bytes_per_iteration_per_gpu": D2H, "inflight": a.inflight, "gpus_visible": have, "runs": []}
    for n in [int(x) for x in a.gpus.split(",")]:
        if n <= have:
            out["runs"].append(run(list(range(n)), a.iters, a.inflight))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
