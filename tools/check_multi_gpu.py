#!/usr/bin/env python
"""N-GPU check of the sharding contract, run under torchrun with the nccl backend:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multi_gpu.py

Every rank produces its shard of the global view ids (view_base = first id), the ranks all-gather the finished views
over NCCL (rlap_b200.dist.all_gather_views), and every rank checks that the gathered rows are bit-identical to all
views computed on one GPU. Prints the all-gather time and the bytes it moved."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlap_b200 import dist as rdist, graphs, ops  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n, total_views, seed = 40000, 4 * world, 77
    ei = torch.from_numpy(graphs.barabasi_albert(n, 7, seed=3)).to(dev)
    g = ops.prepare(ei, None, n)
    ok = True
    for o_v in ("degree", "random", "coarsen"):
        base, cnt = rdist.shard_views(total_views)
        (row, col, w), vp = ops.schur_views(g, n // 2, o_v, "asc", num_views=cnt, seed=seed, view_base=base, dtype=None)
        mine = torch.stack([row.to(torch.float64), col.to(torch.float64), w.to(torch.float64)], dim=1)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        allrows, gvp = rdist.all_gather_views(mine, vp)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        (r1, c1, w1), vp1 = ops.schur_views(g, n // 2, o_v, "asc", num_views=total_views, seed=seed, view_base=0, dtype=None)
        ref = torch.stack([r1.to(torch.float64), c1.to(torch.float64), w1.to(torch.float64)], dim=1)
        same = torch.equal(gvp, vp1) and torch.equal(allrows, ref)
        ok = ok and same
        if rank == 0:
            print(f"{o_v}: {total_views} views over {world} GPUs, gathered {allrows.shape[0]} rows "
                  f"({allrows.numel() * 8 / 1e6:.0f} MB) in {dt * 1e3:.1f} ms, identical to the single-GPU result: {same}",
                  flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-GPU check", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
