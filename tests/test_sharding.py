"""N>1 host logic on CPU: world_size-2 gloo processes shard closed GOPs / streams without overlap and the
timed-region reduction is the max over ranks (what bench.py does under torchrun with NCCL)."""
import os
import sys
import pytest
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


@pytest.mark.gpu
def test_two_devices_in_one_process(oracle, b2):
    """one process may drive several GPUs (kernel attributes and engines are per device): the same GOP encoded on device 0
    and on device 1 gives identical results.  Skipped on single-GPU boxes."""
    import numpy as np
    if b2.lib().b2_device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h = 208, 160
    frames = [oracle.synth_frame(w, h, t, 1) for t in range(3)]
    outs = []
    for dev in (0, 1):
        eng = b2.Engine(w, h, slots=1, ring=1, merange=32, qp=27, device=dev, deblock=1, transform8x8=1, partitions=2, pack_levels=1)
        res = []
        for t, fr in enumerate(frames):
            eng.put_frame(0, 0, list(fr))
            eng.h2d(); eng.encode(b2.FRAME_I if t == 0 else b2.FRAME_P); eng.d2h(); eng.sync()
            info, coef = eng.results(0)
            res.append((info.copy(), coef.copy(), [p.copy() for p in eng.recon(0)]))
        eng.close()
        outs.append(res)
    for a, b in zip(*outs):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1]["blk"], b[1]["blk"])
        assert all(np.array_equal(x, y) for x, y in zip(a[2], b[2]))
