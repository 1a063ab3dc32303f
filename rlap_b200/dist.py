"""Multi-GPU plumbing: one process per GPU (torch.distributed). Views / graphs / seeds are independent
units, so the data path needs no collective: every rank produces its own slice of view ids. The
only exchange is the optional all-gather of finished views for a DDP consumer."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """(view_base, count) of the contiguous slice of global view ids [0, num_views) owned by `rank`.
    Randomness is keyed on the global view id, so the union over ranks equals a single-GPU run."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def all_gather_views(edge_info: Tensor, view_ptr: Tensor, group=None) -> Tuple[Tensor, Tensor]:
    """Variable-length all-gather: every rank holds the rows of its own views ([rows, C] tensor and a
    host int64 view_ptr); returns the rows of ALL views in global view order and the global view_ptr.
    Works with NCCL (device tensors) and gloo (CPU tensors)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return edge_info, view_ptr
    world = dist.get_world_size(group)
    dev = edge_info.device
    counts = (view_ptr[1:] - view_ptr[:-1]).to(torch.int64)
    meta = torch.tensor([edge_info.shape[0], counts.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    rows = [int(m[0]) for m in metas]
    nviews = [int(m[1]) for m in metas]
    max_rows, max_views = max(rows), max(nviews)
    # per-view row counts
    cpad = torch.zeros(max_views, dtype=torch.int64, device=dev)
    cpad[: counts.numel()] = counts.to(dev)
    call = [torch.zeros_like(cpad) for _ in range(world)]
    dist.all_gather(call, cpad, group=group)
    # rows, padded to the longest slab
    pad = torch.zeros((max_rows,) + tuple(edge_info.shape[1:]), dtype=edge_info.dtype, device=dev)
    pad[: edge_info.shape[0]] = edge_info
    slabs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(slabs, pad, group=group)
    out = torch.cat([slabs[r][: rows[r]] for r in range(world)], dim=0)
    allc = torch.cat([call[r][: nviews[r]] for r in range(world)]).cpu()
    vp = torch.zeros(allc.numel() + 1, dtype=torch.int64)
    vp[1:] = torch.cumsum(allc, 0)
    return out, vp
