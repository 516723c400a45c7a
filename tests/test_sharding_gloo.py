"""N > 1 host logic on CPU: two gloo ranks shard a frame batch by index, each extracts its frames (CPU oracle standing in for
the device path: this image has no GPU), summaries are gathered; the result must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from tc2li_slam_b200.sharding import shard_range, weak_scaling_frames
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert weak_scaling_frames(4096, 3) == (12288, 16384)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import pyoracle as po
    from tc2li_slam_b200.sharding import frame_seed, gather_frame_summaries, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_frames, rank, world)
    ex = po.OracleExtractor(500, 1.2, 4, 20, 7)
    local = []
    for f in range(lo, hi):
        for right in (0, 1):
            img = po.synth_v1(320, 160, frame_seed(f), 16 * right, 0x5bd1e995 * right)
            mono, k, d = ex(img)
            local.append((f, right, len(k), int(d.astype(np.uint64).sum())))
    dist.barrier()
    allres = gather_frame_summaries(local, dist)
    if rank == 0:
        q.put(allres)
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_single_process():
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    from tc2li_slam_b200.sharding import frame_seed
    n_frames = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ex = po.OracleExtractor(500, 1.2, 4, 20, 7)
    want = []
    for f in range(n_frames):
        for right in (0, 1):
            mono, k, d = ex(po.synth_v1(320, 160, frame_seed(f), 16 * right, 0x5bd1e995 * right))
            want.append((f, right, len(k), int(d.astype(np.uint64).sum())))
    assert got == want
