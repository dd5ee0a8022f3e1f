"""Frame-wise sharding of an offline sequence across ranks (BASELINE config 3: 4541 stereo pairs over
1/2/4/8 GPUs).  Frames do not interact on the hot path, so a rank owns a contiguous
range of pairs and there is NO collective on the data path; torch.distributed is used only for the
plumbing around it (barrier, max-over-ranks timing, gathering per-rank result digests)."""
import hashlib

import numpy as np


def shard_range(n_total, rank, world):
    """contiguous chunk of ceil(n_total/world) items for `rank` (the last ranks may get fewer/none)"""
    per = -(-n_total // world)
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def digest(arrays):
    """order-sensitive digest of per-frame outputs (keypoints, descriptors, stereo coords)"""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def gather_objects(obj, dist=None):
    """all ranks' python objects, in rank order (single process: [obj])"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(x, dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def process_sequence(extract_pair, n_total, rank, world):
    """runs extract_pair(i) -> list of arrays for every pair index of this rank's shard; returns
    {pair index: digest}.  Concatenating the ranks' dicts reproduces the single-process result."""
    lo, hi = shard_range(n_total, rank, world)
    return {i: digest(extract_pair(i)) for i in range(lo, hi)}
