"""N > 1 host logic on CPU: view sharding and the variable-length all-gather (gloo, world_size 2)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlap_b200 import dist as rdist


def test_shard_views_partitions_the_view_ids():
    for total in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                b, c = rdist.shard_views(total, r, world)
                got += list(range(b, b + c))
            assert got == list(range(total))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, cnt = rdist.shard_views(5, rank, world)
    # fake views: view v has v + 1 rows, every row = (v, v, v)
    rows = [torch.full((v + 1, 3), float(v), dtype=torch.float64) for v in range(base, base + cnt)]
    info = torch.cat(rows) if rows else torch.zeros((0, 3), dtype=torch.float64)
    vp = torch.zeros(cnt + 1, dtype=torch.int64)
    vp[1:] = torch.cumsum(torch.tensor([r.shape[0] for r in rows], dtype=torch.int64), 0)
    out, gvp = rdist.all_gather_views(info, vp)
    q.put((rank, out, gvp))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_views_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_vp = torch.tensor([0, 1, 3, 6, 10, 15])
    want = torch.cat([torch.full((v + 1, 3), float(v), dtype=torch.float64) for v in range(5)])
    for rank, out, gvp in res:
        assert torch.equal(gvp, want_vp)
        assert torch.equal(out, want)
