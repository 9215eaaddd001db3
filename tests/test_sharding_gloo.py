"""CPU: the N>1 host logic (stream partition, final host gather, max-over-ranks timing) with
world_size 2 on the gloo backend."""
import os

import pytest

from pygofpid_b200.sharding import streams_for_rank


def test_partition_is_exact_and_balanced():
    for n in (1, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [streams_for_rank(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert streams_for_rank(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        streams_for_rank(4, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from pygofpid_b200.sharding import gather_detections, max_over_ranks, streams_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = streams_for_rank(5, rank, world)
    local = {s: [("blob", s, rank)] for s in range(lo, hi)}
    merged = gather_detections(local, dist)
    tmax = max_over_ranks([1.0 + rank, 10.0 - rank], dist)
    q.put((rank, merged, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_timing_over_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, merged, tmax = q.get(timeout=120)
        res[rank] = (merged, tmax)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1][0] is None
    assert sorted(res[0][0]) == [0, 1, 2, 3, 4]
    assert res[0][0][4] == [("blob", 4, 1)] and res[0][0][0] == [("blob", 0, 0)]
    assert res[0][1] == [2.0, 10.0] and res[1][1] == [2.0, 10.0]
