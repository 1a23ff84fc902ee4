"""Stream sharding host logic, incl. the N>1 summary gather over gloo with world_size 2 (CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_contiguously():
    from livevideo_b200.shard import max_shard, owner_of, shard_range
    for n in (0, 1, 2, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for a, b in zip(blocks, blocks[1:]):
                assert a[0] + a[1] == b[0]
            assert blocks[-1][0] + blocks[-1][1] == n
            sizes = [c for _, c in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == max_shard(n, world)
            for s in range(n):
                g = owner_of(s, n, world)
                assert blocks[g][0] <= s < blocks[g][0] + blocks[g][1]
    assert shard_range(256, 8, 3) == (96, 32)        # GPU g owns [g*256/G, (g+1)*256/G)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_gather_single_process_is_identity():
    from livevideo_b200.shard import gather_summaries
    s = torch.arange(24, dtype=torch.int32)
    assert torch.equal(gather_summaries(s, 3, 4), s.reshape(3, 4, 2))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, n_frames, w, h, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lvoracle as orc
    from livevideo_b200.shard import gather_summaries, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard_range(n_streams, world, rank)
        # each rank computes the summaries of ITS streams only (here with the CPU oracle standing in for the GPU)
        frames = orc.gen_batch("camera", 1, count, n_frames, w, h, first_stream=first)
        prev = np.zeros((max(count, 1), w * h), dtype=np.uint8)
        r = orc.convert_diff_batch(frames, prev, count, n_frames, w, h, 20, 0, want_bgr=False, want_mask=False)
        local = torch.from_numpy(r["summary"].astype(np.int64).astype(np.int32).reshape(-1))
        full = gather_summaries(local, n_streams, n_frames)
        from livevideo_b200.shard import gather_summaries_async
        pend = [gather_summaries_async(local, n_streams, n_frames) for _ in range(3)]   # several in flight
        for p in pend:
            assert torch.equal(p.result(), full)
        # the peer-memory exchange cannot be set up without a GPU: every rank must learn that together (no rank may
        # hang in the handle swap) and stay on the all-gather
        from livevideo_b200.shard import ShardedContext
        sc = ShardedContext.__new__(ShardedContext)
        sc.total_streams, sc.rank, sc.world, sc.device = n_streams, rank, world, 0
        sc.exchange, sc.exchange_error = None, None
        if not torch.cuda.is_available():
            assert sc.enable_peer_exchange(n_frames) is False
            assert sc.exchange is None and "LiveVideoError" in sc.exchange_error
        q.put((rank, full.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [4, 5])
def test_world2_gather_equals_single_rank(n_streams):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lvoracle as orc
    n_frames, w, h = 3, 64, 48
    frames = orc.gen_batch("camera", 1, n_streams, n_frames, w, h)
    prev = np.zeros((n_streams, w * h), dtype=np.uint8)
    want = orc.convert_diff_batch(frames, prev, n_streams, n_frames, w, h, 20, 0, want_bgr=False,
                                  want_mask=False)["summary"].reshape(n_streams, n_frames, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, n_frames, w, h, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert np.array_equal(got[r].astype(np.uint32), want)   # every rank holds the whole table
