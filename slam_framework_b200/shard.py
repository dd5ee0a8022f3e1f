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
    another: frames do not interact on this path.  The handles (arenas, pinned result buffers) are created by the first run and
    kept until close().

        runner = SequenceRunner(lib, devices=[0, 1, 2, 3], params=dict(nfeatures=2000), batch_pairs=64)
        digests = runner.run(get_pair, n_pairs, bf, baseline)          # {pair index: sha256 of its outputs}

    get_pair(i) -> (left, right) u8 arrays; or get_batch(s, e) -> one contiguous (2 (e - s), H, W) u8 array holding the pairs
    s .. e-1 as L, R, L, R, ... (ideally in pinned memory: it is handed to the DMA engine as it is).  The digests equal those of
    the per-rank run (process_sequence) frame by frame; digests=False returns (n_left, n_right, n_stereo) per pair instead."""

    def __init__(self, lib, devices, params=None, batch_pairs=64, lanes=2):
        self.L, self.devices, self.params, self.B, self.lanes = lib, list(devices), dict(params or {}), int(batch_pairs), int(lanes)
        self.seconds = None
        self._lanes = {}   # device slot -> [[extractor, host buffers, pending pair range], ...]

    def _buffers(self, ex, n_img):
        """make_buffers layout in PORTABLE pinned host memory (orbfe_pinned_alloc): the D2H copies need no staging, on any device"""
        from . import orbfe
        try:
            cap = ex.max_keypoints()
            mk = lambda shape, dt: orbfe.pinned_empty(shape, dt, lib=self.L)
            return dict(kps=mk((n_img, cap), orbfe.KP_DTYPE), desc=mk((n_img, cap, 32), np.uint8), n=mk((n_img,), np.int32), cap=cap,
                        ur=mk((n_img, cap), np.float32), depth=mk((n_img, cap), np.float32))
        except Exception:
            return ex.make_buffers(n_img, stereo=True)

    def _device_loop(self, slot, dev, lo, hi, get_pair, get_batch, bf, baseline, out, errors, digests):
        import ctypes
        from . import orbfe
        try:
            if get_batch is not None:
                h, w = get_batch(lo, lo + 1).shape[1:]
            else:
                h, w = get_pair(lo)[0].shape
            lanes = self._lanes.get(slot)
            if lanes is None or lanes[0][0].max_images != 2 * self.B:
                lanes = []
                for _ in range(self.lanes):
                    ex = orbfe.ORBextractor(device=dev, max_images=2 * self.B, max_size=(w, h), lib=self.L, **self.params)
                    lanes.append([ex, self._buffers(ex, 2 * self.B), None])   # handle, host buffers, pending pair range
                self._lanes[slot] = lanes

            def collect(lane):
                ex, buf, pend = lane[0], lane[1], lane[2]
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
                ex, buf = lane[0], lane[1]
                if get_batch is not None:
                    frames = get_batch(s, e)
                    n_img = len(frames)
                    ptrs = (ctypes.c_void_p * n_img)(*[frames[i].ctypes.data for i in range(n_img)])
                    ex.upload_ptrs(ptrs, n_img, w, h, frames.strides[1])
                    lane.append(frames)            # keeps the source alive until the copy has been consumed
                    del lane[3:-1]
                else:
                    imgs = [im for i in range(s, e) for im in get_pair(i)]
                    n_img = len(imgs)
                    ex.upload(imgs)
                ex.run(n_img)
                ex.run_stereo(e - s, bf, baseline)
                ex.download_async(n_img, buf)
                lane[2] = (s, e)
            for lane in lanes:
                collect(lane)
                del lane[3:]
        except Exception as exc:  # surfaced by run()
            errors.append((slot, dev, exc))

    def run(self, get_pair, n_pairs, bf, baseline, digests=True, get_batch=None):
        import threading
        import time
        out, errors, threads = {}, [], []
        t0 = time.perf_counter()
        for slot, dev in enumerate(self.devices):
            lo, hi = shard_range(n_pairs, slot, len(self.devices))
            if hi <= lo:
                continue
            t = threading.Thread(target=self._device_loop, args=(slot, dev, lo, hi, get_pair, get_batch, bf, baseline, out, errors, digests))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        self.seconds = time.perf_counter() - t0
        if errors:
            raise RuntimeError(f"device loop failed: {errors[0]}")
        return out

    def close(self):
        for lanes in self._lanes.values():
            for lane in lanes:
                lane[0].close()
        self._lanes = {}
