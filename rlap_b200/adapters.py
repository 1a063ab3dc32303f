"""Plug-in points of the reference's callers, without PyGCL / PyG / DGL:

 * `rLap` — the augmentor of scripts/augmentor_benchmarks.py:68-96 (PyGCL `Augmentor` protocol: `augment(Graph)`
   and `__call__(x, edge_index, edge_weight)`), and of CCA-SSG/aug.py:33-72 (edge list in, edge list out);
   the graph stays on the device, node ids are int64 (the reference round-trips them through float32).
 * `rLapPPRDiffusion` — scripts/augmentor_benchmarks.py:99-171: Schur view -> relabelled subgraph -> PPR diffusion.
 * `compact_relabel` — torch.unique(sorted) + subgraph(relabel_nodes=True) as used at :149-155.
 * `chained_schur_stats` — scripts/rlap_vc_spectral.py:14-58: repeated elimination with per-step node count,
   edge count and top singular value of the unweighted adjacency.
"""
from collections import namedtuple
from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor

from . import ops


class Graph(namedtuple("Graph", ["x", "edge_index", "edge_weights"])):
    """the (x, edge_index, edge_weights) triple PyGCL augmentors pass around"""

    def unfold(self):
        return self.x, self.edge_index, self.edge_weights


def _num_nodes(edge_index: Tensor) -> int:
    return int(edge_index.max().item()) + 1 if edge_index.numel() else 0


class rLap:
    """randomized Schur-complement augmentor: eliminates int(frac * num_nodes) nodes (num_nodes =
    edge_index.max() + 1, augmentor_benchmarks.py:77-78) and returns the sampled edge list; weights are dropped
    unless keep_weights=True (the reference returns edge_weights=None, :96)."""

    def __init__(self, frac: float, o_v: str = "random", o_n: str = "asc", keep_weights: bool = False,
                 seed: Optional[int] = None):
        self.frac, self.o_v, self.o_n = frac, o_v, o_n
        self.keep_weights = keep_weights
        self.seed = seed
        self.num_remove = 0
        self._views_drawn = 0     # a seeded augmentor draws views seed/0, seed/1, ...: fresh randomness at every call,
                                  # reproducible as a whole (the reference draws from std::random_device every time)

    def views(self, edge_index: Tensor, edge_weights: Optional[Tensor] = None, num_views: int = 1,
              num_nodes: Optional[int] = None, relabel: bool = False):
        """num_views independent views in one batched call: [(edge_index [2,E'] int64, weights or None), ...].
        relabel=True: the views come compacted to their surviving nodes (relabelled inside the emission kernel) and
        every entry is (nodes, edge_index', weights or None) like `compact_relabel` returns."""
        n = _num_nodes(edge_index) if num_nodes is None else num_nodes
        self.num_remove = int(self.frac * n)
        g = ops.prepare(edge_index, edge_weights, n)
        base = self._views_drawn
        self._views_drawn += num_views
        res = ops.schur_views(g, self.num_remove, self.o_v, self.o_n, num_views=num_views, seed=self.seed,
                              view_base=base if self.seed is not None else 0, dtype=None, weights=self.keep_weights,
                              relabel=relabel)
        ((row, col, w), newid), vp = (res[0] if relabel else (res[0], None)), res[1]
        out = []
        for v in range(num_views):
            s, e = int(vp[v]), int(vp[v + 1])
            ei = torch.stack([row[s:e], col[s:e]]).long().to(edge_index.device)
            wv = w[s:e].to(edge_index.device) if self.keep_weights else None
            if relabel:
                out.append(((newid[v] >= 0).nonzero().reshape(-1).to(edge_index.device), ei, wv))
            else:
                out.append((ei, wv))
        return out

    def augment(self, g):
        x, edge_index, edge_weights = g.unfold()
        ei, w = self.views(edge_index, edge_weights, 1)[0]
        return type(g)(x=x, edge_index=ei, edge_weights=w) if hasattr(g, "_replace") else Graph(x, ei, w)

    def __call__(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None):
        return self.augment(Graph(x, edge_index, edge_weight)).unfold()


def compact_relabel(edge_index: Tensor, edge_weights: Optional[Tensor] = None):
    """nodes = sorted unique endpoints; edges relabelled to 0..len(nodes)-1. Returns (nodes, edge_index', weights)."""
    nodes = torch.unique(edge_index, sorted=True)
    return nodes, torch.searchsorted(nodes, edge_index), edge_weights


def ppr_dense(edge_index: Tensor, edge_weight: Optional[Tensor], alpha: float = 0.2, eps: float = 1e-4):
    """exact personalised-PageRank diffusion of a small graph, as torch_geometric's GDC does it (sym-normalised
    transition matrix T, alpha (I - (1 - alpha) T)^-1, threshold eps, sym-normalise again). Dense: for the
    relabelled Schur subgraphs of node-level datasets."""
    n = _num_nodes(edge_index)
    w = torch.ones(edge_index.shape[1], device=edge_index.device) if edge_weight is None else edge_weight.float()
    A = torch.zeros((n, n), device=edge_index.device)
    A.index_put_((edge_index[0], edge_index[1]), w, accumulate=True)

    def sym(M):
        d = M.sum(1)
        dis = torch.where(d > 0, d.pow(-0.5), torch.zeros_like(d))
        return dis[:, None] * M * dis[None, :]

    T = sym(A)
    S = alpha * torch.linalg.inv(torch.eye(n, device=A.device) - (1 - alpha) * T)
    S = torch.where(S >= eps, S, torch.zeros_like(S))
    S = sym(S)
    idx = S.nonzero(as_tuple=False).t()
    return idx, S[idx[0], idx[1]]


class rLapPPRDiffusion:
    """Schur view, relabelled to its surviving nodes, then diffused (augmentor_benchmarks.py:99-171). `diffusion`
    is any callable (edge_index, edge_weight) -> (edge_index, edge_weight), e.g. PyGCL's compute_ppr; the default
    is ppr_dense above. The cache / refresh_cache_freq behaviour of the reference is kept."""

    def __init__(self, frac, o_v="random", o_n="asc", alpha=0.2, eps=1e-4, use_cache=True, refresh_cache_freq=50,
                 diffusion: Optional[Callable] = None, seed: Optional[int] = None):
        self.rlap = rLap(frac, o_v, o_n, keep_weights=True, seed=seed)
        self.alpha, self.eps = alpha, eps
        self.diffusion = diffusion
        self._cache = None
        self.use_cache = use_cache
        self.refresh_cache_freq = refresh_cache_freq
        self.refresh_cache_counter = 0

    def augment(self, g):
        if self._cache is not None and self.use_cache and self.refresh_cache_counter < self.refresh_cache_freq:
            self.refresh_cache_counter += 1
            return self._cache
        x, edge_index, edge_weights = g.unfold()
        nodes, sub_ei, sub_w = self.rlap.views(edge_index, edge_weights, 1, relabel=True)[0]   # compacted by the emission kernel
        if self.diffusion is not None:
            dei, dw = self.diffusion(sub_ei, sub_w)
        else:
            dei, dw = ppr_dense(sub_ei, sub_w, alpha=self.alpha, eps=self.eps)
        res = Graph(x=x, edge_index=nodes[dei], edge_weights=dw)
        self._cache = res
        self.refresh_cache_counter = 0
        return res

    def __call__(self, x, edge_index, edge_weight=None):
        return self.augment(Graph(x, edge_index, edge_weight)).unfold()


def chained_schur_stats(edge_index: Tensor, edge_weights: Optional[Tensor], batch_count: int, nodes_to_eliminate: int,
                        o_v: str, o_n: str, seed: Optional[int] = None, approximate=None):
    """scripts/rlap_vc_spectral.py:14-58: `batch_count` successive eliminations of `nodes_to_eliminate` nodes, each on
    the relabelled output of the previous one (weights carried). Returns (max_sv, num_unique_nodes, num_edges) per
    step; max_sv is the top singular value of the dense unweighted adjacency (exact spectral norm here, the
    reference estimates it with torch.svd_lowrank). `approximate` lets tests swap in another implementation of
    approximate_cholesky(edge_index, edge_weights, num_nodes, num_remove, o_v, o_n) -> [E',3]."""
    fn = approximate or (lambda ei, ew, n, t, ov, on: ops.approximate_cholesky(ei, ew, n, t, ov, on, seed=None if seed is None else seed + len(max_sv)))
    max_sv, num_unique_nodes, num_edges = [], [], []
    num_nodes = _num_nodes(edge_index)
    for _ in range(batch_count):
        info = fn(edge_index, edge_weights, num_nodes, nodes_to_eliminate, o_v, o_n)
        ei = info[:, :2].long().t().contiguous()
        nodes, ei, w = compact_relabel(ei, info[:, 2])
        num_unique_nodes.append(int(nodes.shape[0]))
        num_nodes = int(nodes.shape[0])
        edge_index, edge_weights = ei, w
        num_edges.append(int(ei.shape[1]))
        adj = torch.zeros((num_nodes, num_nodes), device=ei.device)
        adj[ei[0], ei[1]] = 1.0
        max_sv.append(float(torch.linalg.matrix_norm(adj, ord=2)))
    return max_sv, num_unique_nodes, num_edges
