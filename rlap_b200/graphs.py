"""Seeded synthetic graph generators for the BASELINE.json configs.

No datasets / torch_geometric in this image, so the shapes the reference's call sites
use are restated in numpy: the PyG-style Barabasi-Albert generator used by
tests/test_rlap.py:27-30 (+ to_undirected) and a stochastic block model.
All functions return a directed, symmetric, duplicate-free edge_index [2,E] int64.
"""
import numpy as np


def _symmetrize(row: np.ndarray, col: np.ndarray, n: int) -> np.ndarray:
    keep = row != col
    row, col = row[keep], col[keep]
    r = np.concatenate([row, col]).astype(np.int64)
    c = np.concatenate([col, row]).astype(np.int64)
    key = np.unique(r * n + c)
    return np.stack([key // n, key % n])


def barabasi_albert(n: int, m: int, seed: int = 0) -> np.ndarray:
    """PyG barabasi_albert_graph(n, m) followed by to_undirected (tests/test_rlap.py:27-30):
    start with arange(m) <-> randperm(m); every new node draws m endpoints with replacement
    from the current endpoint multiset; symmetrise and dedupe."""
    assert 0 < m < n
    rng = np.random.default_rng(seed)
    cap = 2 * m + 2 * m * (n - m)
    ends = np.empty(cap, dtype=np.int64)          # endpoint multiset (row ++ col of PyG)
    row = np.empty(m + m * (n - m), dtype=np.int64)
    col = np.empty_like(row)
    row[:m] = np.arange(m)
    col[:m] = rng.permutation(m)
    ends[:m] = row[:m]
    ends[m:2 * m] = col[:m]
    ne, nr = 2 * m, m
    for i in range(m, n):
        pick = ends[rng.integers(0, ne, size=m)]
        row[nr:nr + m] = pick
        col[nr:nr + m] = i
        ends[ne:ne + m] = pick
        ends[ne + m:ne + 2 * m] = i
        ne += 2 * m
        nr += m
    return _symmetrize(row[:nr], col[:nr], n)


def sbm(n: int, n_blocks: int, n_undirected: int, p_ratio: float = 10.0, seed: int = 0) -> np.ndarray:
    """Stochastic block model with n_blocks near-equal blocks and EXACTLY n_undirected
    undirected edges: pairs are drawn with within-block probability p_ratio times the
    cross-block one and resampled until the count is hit."""
    rng = np.random.default_rng(seed)
    block = (np.arange(n) * n_blocks) // n
    sizes = np.bincount(block, minlength=n_blocks)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    pairs_in = float(np.sum(sizes * (sizes - 1) // 2))
    pairs_all = n * (n - 1) / 2.0
    w_in = p_ratio * pairs_in
    frac_in = w_in / (w_in + (pairs_all - pairs_in))
    keys = np.empty(0, dtype=np.int64)
    while keys.size < n_undirected:
        need = n_undirected - keys.size
        k = int(need * 1.2) + 16
        n_in = rng.binomial(k, frac_in)
        # within-block pairs: pick a block proportional to its pair count, then two members
        pb = sizes * (sizes - 1) / 2.0
        b = rng.choice(n_blocks, size=n_in, p=pb / pb.sum())
        u = starts[b] + (rng.random(n_in) * sizes[b]).astype(np.int64)
        v = starts[b] + (rng.random(n_in) * sizes[b]).astype(np.int64)
        # cross-block pairs: rejection on block equality
        uo = rng.integers(0, n, size=k - n_in)
        vo = rng.integers(0, n, size=k - n_in)
        ok = block[uo] != block[vo]
        u = np.concatenate([u, uo[ok]])
        v = np.concatenate([v, vo[ok]])
        ok = u != v
        lo, hi = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
        new = np.unique(lo * n + hi)
        new = np.setdiff1d(new, keys, assume_unique=True)
        if new.size > need:
            new = rng.permutation(new)[:need]
        keys = np.union1d(keys, new)
    lo, hi = keys // n, keys % n
    return _symmetrize(lo, hi, n)


def proteins_like_batch(n_graphs: int = 1113, mean_nodes: float = 39.06, seed: int = 0):
    """PROTEINS-shaped batch (C3): node counts from a clipped lognormal rescaled to the
    PROTEINS_full mean, each graph BA(m=2). Returns (edge_index of the disjoint union,
    graph_ptr [n_graphs+1]) - the union is what a PyG Batch hands to the augmentor
    (scripts/graph_shared.py:139-146)."""
    rng = np.random.default_rng(seed)
    raw = rng.lognormal(mean=3.3, sigma=0.7, size=n_graphs)
    sizes = np.clip(raw * (mean_nodes / raw.mean()), 4, 620).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    parts = []
    for g in range(n_graphs):
        ei = barabasi_albert(int(sizes[g]), 2, seed=seed * 1000003 + g)
        parts.append(ei + ptr[g])
    return np.concatenate(parts, axis=1), ptr


def sbm_torch(n: int, n_blocks: int, n_undirected: int, p_ratio: float = 10.0, seed: int = 0, device="cuda"):
    """The same stochastic block model as `sbm`, generated with torch on `device` (the ogbn-products-shaped graph of
    C5 has 61.9 M undirected edges: seconds on a GPU, minutes in numpy). Returns a torch int64 edge_index [2, 2 *
    n_undirected] on `device`: directed, symmetric, duplicate free, sorted by (row, col). Not the same edges as `sbm`
    for the same seed (different random streams)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    dev = torch.device(device)
    block = (torch.arange(n, device=dev, dtype=torch.int64) * n_blocks) // n
    sizes = torch.bincount(block, minlength=n_blocks)
    starts = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(sizes, 0)])
    pb = (sizes * (sizes - 1) / 2.0).double()
    pairs_in = float(pb.sum())
    pairs_all = n * (n - 1) / 2.0
    w_in = p_ratio * pairs_in
    frac_in = w_in / (w_in + (pairs_all - pairs_in))
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    while keys.numel() < n_undirected:
        need = n_undirected - keys.numel()
        k = int(need * 1.2) + 16
        n_in = int(round(k * frac_in))
        b = torch.multinomial((pb / pb.sum()).float(), n_in, replacement=True, generator=gen)
        u = starts[b] + (torch.rand(n_in, device=dev, generator=gen, dtype=torch.float64) * sizes[b]).long()
        v = starts[b] + (torch.rand(n_in, device=dev, generator=gen, dtype=torch.float64) * sizes[b]).long()
        uo = torch.randint(0, n, (k - n_in,), device=dev, generator=gen)
        vo = torch.randint(0, n, (k - n_in,), device=dev, generator=gen)
        ok = block[uo] != block[vo]
        u = torch.cat([u, uo[ok]])
        v = torch.cat([v, vo[ok]])
        ok = u != v
        lo, hi = torch.minimum(u[ok], v[ok]), torch.maximum(u[ok], v[ok])
        new = torch.unique(lo * n + hi)
        if keys.numel():
            new = new[~torch.isin(new, keys)]
        if new.numel() > need:
            new = new[torch.randperm(new.numel(), device=dev, generator=gen)[:need]]
        keys = torch.unique(torch.cat([keys, new]))
    lo, hi = keys // n, keys % n
    both = torch.sort(torch.cat([lo * n + hi, hi * n + lo])).values
    return torch.stack([both // n, both % n])
