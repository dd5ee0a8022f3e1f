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


class SequenceRunner:
    """Single-process multi-GPU driver of an offline sequence (the reference is ONE process: examples/main_stereo.cpp:102-143
    walks the sequence in a loop).  The pairs are sharded contiguously across `devices` (shard_range); each device gets one host
    thread (the C-ABI calls release the GIL) and `lanes` extractor handles = `lanes` streams, which take the device's batches in
    turn so that the H2D copy of batch k+1 and the D2H copy of batch k-1 overlap the kernels of batch k.  No device talks to
    another: frames do not interact on this path.

        runner = SequenceRunner(lib, devices=[0, 1, 2, 3], params=dict(nfeatures=2000), batch_pairs=64)
        digests = runner.run(get_pair, n_pairs, bf, baseline)          # {pair index: sha256 of its outputs}

    get_pair(i) -> (left, right) u8 arrays.  The digests equal those of the per-rank run (process_sequence) frame by frame."""

    def __init__(self, lib, devices, params=None, batch_pairs=64, lanes=2):
        self.L, self.devices, self.params, self.B, self.lanes = lib, list(devices), dict(params or {}), int(batch_pairs), int(lanes)
        self.seconds = None

    def _device_loop(self, slot, dev, lo, hi, get_pair, bf, baseline, out, errors, digests=True):
        from . import orbfe
        try:
            first = get_pair(lo)[0]
            h, w = first.shape
            lanes = []
            for _ in range(self.lanes):
                ex = orbfe.ORBextractor(device=dev, max_images=2 * self.B, max_size=(w, h), lib=self.L, **self.params)
                lanes.append([ex, ex.make_buffers(2 * self.B, stereo=True), None])   # handle, host buffers, pending pair range

            def collect(lane):
                ex, buf, pend = lane
                if pend is None:
                    return
                ex.sync()
                for k, i in enumerate(range(*pend)):
                    n0, n1 = int(buf["n"][2 * k]), int(buf["n"][2 * k + 1])
                    if not digests:   # throughput runs: keypoint counts and stereo matches only
                        out[i] = (n0, n1, int((buf["ur"][2 * k, :n0] >= 0).sum()))
                        continue
                    out[i] = digest([buf["kps"][2 * k, :n0], buf["desc"][2 * k, :n0], buf["kps"][2 * k + 1, :n1], buf["desc"][2 * k + 1, :n1],
                                     buf["ur"][2 * k, :n0], buf["depth"][2 * k, :n0]])
                lane[2] = None
            for j, s in enumerate(range(lo, hi, self.B)):
                lane = lanes[j % self.lanes]
                collect(lane)                      # the batch this handle took `lanes` batches ago has delivered
                e = min(s + self.B, hi)
                imgs = [im for i in range(s, e) for im in get_pair(i)]
                ex, buf, _ = lane
                ex.upload(imgs)
                ex.run(len(imgs))
                ex.run_stereo(e - s, bf, baseline)
                ex.download_async(len(imgs), buf)
                lane[2] = (s, e)
            for lane in lanes:
                collect(lane)
            for ex, _, _ in lanes:
                ex.close()
        except Exception as exc:  # surfaced by run()
            errors.append((slot, dev, exc))

    def run(self, get_pair, n_pairs, bf, baseline, digests=True):
        import threading
        import time
        out, errors, threads = {}, [], []
        t0 = time.perf_counter()
        for slot, dev in enumerate(self.devices):
            lo, hi = shard_range(n_pairs, slot, len(self.devices))
            if hi <= lo:
                continue
            t = threading.Thread(target=self._device_loop, args=(slot, dev, lo, hi, get_pair, bf, baseline, out, errors, digests))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        self.seconds = time.perf_counter() - t0
        if errors:
            raise RuntimeError(f"device loop failed: {errors[0]}")
        return out
