"""N>1 host logic on CPU: world_size-2 gloo processes shard closed GOPs / streams without overlap and the
timed-region reduction is the max over ranks (what bench.py does under torchrun with NCCL)."""
import os
import sys
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200"))
    import torch
    import torch.distributed as dist
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.units_of_rank(37, rank, world)
    assert all(sharding.owner(u, world) == rank for u in mine)
    t = torch.zeros(37, dtype=torch.int64)
    t[mine] = 1
    dist.all_reduce(t)                                   # every unit encoded exactly once
    assert bool((t == 1).all())
    streams = torch.tensor(sharding.slot_streams(rank, 4))
    allst = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allst, streams)
    assert len(set(torch.cat(allst).tolist())) == 4 * world
    dist.barrier()
    m = sharding.max_over_ranks(1.0 + rank, dist)
    assert m == float(world)
    out.put((rank, len(mine)))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get() for _ in range(2))
    assert got == {0: 19, 1: 18}


def test_single_rank_helpers():
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200"))
    import sharding
    assert sharding.units_of_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks(3.5) == 3.5
