"""world_size-2 gloo test (CPU) of the N>1 path: video sharding + token-id gather in global order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vcb200  # noqa: F401
from vcb200.sharding import IdGatherer, gather_ids, shard_range


def test_shard_range_covers_all_videos_once():
    for n in (0, 1, 5, 64, 65, 512):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_ids(v: int, max_new: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + v)       # what "captioning video v" yields, independent of the shard
    return torch.randint(0, 50257, (max_new,), generator=g, dtype=torch.int32)


def _worker(rank, world, port, n_videos, max_new, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_videos, world, rank)
    ids = torch.stack([_fake_ids(v, max_new) for v in range(lo, hi)]) if hi > lo else torch.zeros(0, max_new, dtype=torch.int32)
    lens = torch.tensor([(v % max_new) + 1 for v in range(lo, hi)], dtype=torch.int32)
    all_ids, all_len = gather_ids(ids, lens, n_videos)
    # the preallocated single-collective gatherer bench.py uses: same global order, own block verified
    per = (n_videos + world - 1) // world
    g = IdGatherer(per, max_new, world, "cpu")
    buf = g.gather(ids, lens)
    assert g.check_own_block(buf, ids, lens, rank)
    g_ids, g_len = g.global_order(buf, n_videos)
    assert torch.equal(g_ids, all_ids) and torch.equal(g_len, all_len)
    buf2 = g.gather(ids, lens)                          # reused buffers: a second batch gives the same answer
    assert buf2.data_ptr() == buf.data_ptr() and g.check_own_block(buf2, ids, lens, rank)
    q.put((rank, all_ids.tolist(), all_len.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [8, 5])          # even split and ragged split
def test_gather_ids_world2_gloo(n_videos):
    world, max_new = 2, 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_videos, max_new, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_ids = [_fake_ids(v, max_new).tolist() for v in range(n_videos)]
    want_len = [(v % max_new) + 1 for v in range(n_videos)]
    for rank, ids, lens in got:
        assert ids == want_ids and lens == want_len
