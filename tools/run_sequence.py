#!/usr/bin/env python
"""The offline sequence (BASELINE.json configs[2], 4541 synthetic KITTI-size stereo pairs) through shard.SequenceRunner: ONE
process driving every visible GPU (a host thread and `--lanes` streams per device), host arrays in, host arrays out.
Prints one JSON line with the wall-clock pairs/s, for comparison with bench.py's per-rank (torchrun) end-to-end figure.

    python tools/run_sequence.py [--gpus N] [--pairs 4541] [--batch 64] [--lanes 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--pairs", type=int, default=B.SEQ_PAIRS)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--distinct", type=int, default=64)
    ap.add_argument("--digests", action="store_true", help="sha256 every pair's outputs (the parity form; slow in python)")
    a = ap.parse_args()
    from slam_framework_b200 import orbfe, shard
    L = orbfe.load()
    n = a.gpus or L.orbfe_device_count()
    pairs = B.make_pairs(a.distinct, 0)
    import numpy as np
    # the sequence's frames in portable pinned memory, as a capture / decode thread would deliver them.  Pair i of the sequence is
    # synthetic pair i % distinct; the block holds them twice over so that every batch [s, e) is ONE contiguous window of it
    # (a wrap-around batch assembled with np.concatenate would be a 60 MB pageable copy per batch: that, not the GPUs, is what
    # an earlier version of this tool measured on shards that do not start at a multiple of `distinct`)
    assert a.batch <= a.distinct
    tiled = orbfe.pinned_empty((4 * a.distinct, B.H, B.W), np.uint8, lib=L)
    for p_ in range(2 * a.distinct):
        tiled[2 * p_], tiled[2 * p_ + 1] = pairs[p_ % a.distinct]

    def get_batch(s, e):
        s0 = s % a.distinct
        return tiled[2 * s0:2 * (s0 + e - s)]
    runner = shard.SequenceRunner(L, devices=list(range(n)), params=dict(nfeatures=B.NFEATURES, scaleFactor=B.SCALE, nlevels=B.NLEVELS,
                                                                          iniThFAST=B.INI_TH, minThFAST=B.MIN_TH),
                                  batch_pairs=a.batch, lanes=a.lanes)
    gb = None if a.digests else get_batch
    runner.run(lambda i: pairs[i % a.distinct], min(a.pairs, 8 * a.batch * n), B.BF, B.BF / B.FX, digests=False, get_batch=gb)   # warm-up: arenas, clocks
    d = runner.run(lambda i: pairs[i % a.distinct], a.pairs, B.BF, B.BF / B.FX, digests=a.digests, get_batch=gb)
    runner.close()
    print(json.dumps({"driver": "single process, one host thread per GPU", "gpus": n, "pairs": a.pairs, "batch_pairs": a.batch, "lanes": a.lanes,
                      "seconds": runner.seconds, "pairs_per_s": a.pairs / runner.seconds, "results": len(d),
                      "digests": a.digests, "note": "wall clock of the second pass (handles kept from the first); frames and results in pinned host memory; python driver"}))


if __name__ == "__main__":
    main()
