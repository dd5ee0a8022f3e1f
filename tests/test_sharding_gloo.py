"""N>1 path on CPU: world_size-2 gloo processes shard a short synthetic sequence frame-wise; the
per-frame digests gathered from the ranks must equal the single-process run (frames do not interact,
so there is no collective on the data path).  Compute runs on the emulated TEST build here (no GPU in
this container); on the GPU box bench.py runs the same sharding with the sm_100a library."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam_framework_b200 import shard

N_PAIRS = 5
HERE = os.path.dirname(os.path.abspath(__file__))


def _pair_fn():
    sys.path.insert(0, HERE)
    from emu import build_emu
    from slam_framework_b200 import orbfe, synth
    L = orbfe.load(build_emu.build(), _test_emulation=True)
    ex = orbfe.ORBextractor(400, lib=L, max_images=2)

    def f(i):
        l, r = synth.stereo_pair(120, 400, seed=50 + i)
        ex.upload([l, r]); ex.run(2); ex.run_stereo(1, 386.1448, 386.1448 / 718.856)
        b = ex.download(2, ex.make_buffers(2, stereo=True))
        n0, n1 = b["n"]
        return [b["kps"][0, :n0], b["desc"][0, :n0], b["kps"][1, :n1], b["desc"][1, :n1], b["ur"][0, :n0], b["depth"][0, :n0]]
    return f


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.process_sequence(_pair_fn(), N_PAIRS, rank, world)
    parts = shard.gather_objects(mine, dist)
    t = shard.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        merged = {}
        for p in parts:
            merged.update(p)
        q.put((merged, t, [sorted(p) for p in parts]))
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for n in (0, 1, 7, 4541):
        for world in (1, 2, 4, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    assert shard.shard_range(4541, 7, 8) == (3976, 4541)


def test_two_rank_gloo_sharding_matches_single_process():
    build_first = _pair_fn()  # builds the emulated library once, before forking workers
    single = shard.process_sequence(build_first, N_PAIRS, 0, 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, t, owned = q.get(timeout=300)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert owned == [[0, 1, 2], [3, 4]]
    assert merged == single
    assert t == 2.0  # max over ranks


def test_single_process_runner_matches_per_rank_run():
    """shard.SequenceRunner (one process, one host thread per device, two handles per device taking the batches in turn) against
    the per-rank run: same digest for every pair.  Two logical devices on the one emulated device: the sharding, the lane
    rotation, partial last batches and the async downloads are what is under test here."""
    sys.path.insert(0, HERE)
    from emu import build_emu
    from slam_framework_b200 import orbfe, synth
    single = shard.process_sequence(_pair_fn(), N_PAIRS, 0, 1)
    L = orbfe.load(build_emu.build(), _test_emulation=True)
    runner = shard.SequenceRunner(L, devices=[0, 0], params=dict(nfeatures=400), batch_pairs=2, lanes=2)
    got = runner.run(lambda i: synth.stereo_pair(120, 400, seed=50 + i), N_PAIRS, 386.1448, 386.1448 / 718.856)
    assert got == single
    assert runner.seconds > 0
    # the batch form (one contiguous block of frames per batch, handed to the copy engine as it is) on the same, kept handles
    import numpy as np
    frames = np.stack([im for i in range(N_PAIRS) for im in synth.stereo_pair(120, 400, seed=50 + i)])
    assert runner.run(None, N_PAIRS, 386.1448, 386.1448 / 718.856, get_batch=lambda s, e: frames[2 * s:2 * e]) == single
    counts = runner.run(None, N_PAIRS, 386.1448, 386.1448 / 718.856, digests=False, get_batch=lambda s, e: frames[2 * s:2 * e])
    assert sorted(counts) == list(range(N_PAIRS)) and all(len(c) == 3 and c[0] > 0 for c in counts.values())
    runner.close()
