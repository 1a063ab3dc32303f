#!/usr/bin/env python
"""Latency / memory harness of the rLap augmentor on synthetic graphs of the reference's dataset shapes.

Mirrors scripts/augmentor_benchmarks.py:366-470 + run_augmentor_benchmarks.sh of the reference:

    python tools/augmentor_benchmarks.py node  rLap CORA        [--repeat 10] [--views V]
    python tools/augmentor_benchmarks.py graph rLap PROTEINS    [--repeat 10]

`node`: one call of the augmentor on the whole graph (fraction 0.5, like the reference); `graph`: the dataset in
batches of 128 graphs, every batch augmented as one union graph (the reference's DataLoader(batch_size=128) loop).
Prints the reference's "DURATION: <sec> sec" line per repeat plus the peak device memory; there is no network,
so the graphs are synthetic with the published node / edge counts of each dataset (SBM for the node datasets,
BA(m) graphs with the TU datasets' mean sizes for the graph datasets). Only the rLap augmentor is built here:
the other rows of the reference's table are PyGCL augmentors, outside the path this repository replaces.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlap_b200 import adapters, graphs  # noqa: E402

# name: (nodes, undirected edges, SBM blocks = classes)
NODE_SHAPES = {"CORA": (2708, 5278, 7), "AMAZON-PHOTO": (7650, 119081, 8), "PUBMED": (19717, 44324, 3),
               "COAUTHOR-CS": (18333, 81894, 15), "COAUTHOR-PHY": (34493, 247962, 5)}
# name: (graphs, mean nodes, BA m ~ mean undirected edges per node)
GRAPH_SHAPES = {"PROTEINS": (1113, 39.06, 2), "MUTAG": (188, 17.93, 1), "IMDB-BINARY": (1000, 19.77, 5),
                "IMDB-MULTI": (1500, 13.0, 5), "NCI1": (4110, 29.87, 1)}


def ba_batch(n_graphs, mean_nodes, m, seed=0):
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.rint(rng.lognormal(np.log(mean_nodes) - 0.125, 0.5, n_graphs)), max(4, m + 2), 620).astype(np.int64)
    parts, ptr = [], [0]
    for i, ng in enumerate(sizes):
        ei = graphs.barabasi_albert(int(ng), m, seed=seed * 100003 + i)
        parts.append(ei + ptr[-1])
        ptr.append(ptr[-1] + int(ng))
    return parts, np.asarray(ptr, dtype=np.int64)


def build_batches(task, dataset, device=None):
    """[(x, edge_index)] as the reference's loops see them: one whole graph (node task) or unions of 128 graphs"""
    mk = (lambda a: torch.from_numpy(a).to(device)) if device is not None else (lambda a: torch.from_numpy(a))
    if task == "node":
        n, e_und, blocks = NODE_SHAPES[dataset]
        ei = mk(graphs.sbm(n, blocks, e_und, seed=0))
        return [(torch.zeros((n, 1), device=ei.device), ei)]
    ng, mean_nodes, m = GRAPH_SHAPES[dataset]
    parts, ptr = ba_batch(ng, mean_nodes, m)
    batches = []
    for b0 in range(0, ng, 128):       # DataLoader(dataset, batch_size=128): a batch is one union graph
        b1 = min(b0 + 128, ng)
        ei = mk(np.concatenate(parts[b0:b1], axis=1) - ptr[b0])
        batches.append((torch.zeros((int(ptr[b1] - ptr[b0]), 1), device=ei.device), ei))
    return batches


def time_augmentor(aug, batches, repeat, sync=None):
    """the reference's timing loop (scripts/augmentor_benchmarks.py:366-393): seconds per pass over the batches"""
    sync = sync or (lambda: None)
    aug(*batches[0], None)                  # build / load the extension, warm the allocator
    sync()
    out_t = []
    for _ in range(repeat):
        duration = 0.0
        for x, ei in batches:
            sync()
            start = time.time()
            aug(x, ei, None)
            sync()
            duration += time.time() - start
        out_t.append(duration)
    return out_t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("task", choices=["node", "graph"])
    ap.add_argument("augmentor", choices=["rLap"])
    ap.add_argument("dataset")
    ap.add_argument("device", nargs="?", default="cuda")
    ap.add_argument("--repeat", type=int, default=10)
    ap.add_argument("--o_v", default="random")
    ap.add_argument("--o_n", default="asc")
    args = ap.parse_args()
    print(args)
    if args.device != "cuda" or not torch.cuda.is_available():
        raise SystemExit("rlap_b200 has no CPU path: run with device 'cuda' on a GPU box")
    dev = torch.device("cuda")
    fraction = 0.5
    aug = adapters.rLap(fraction, o_v=args.o_v, o_n=args.o_n)
    batches = build_batches(args.task, args.dataset, dev)
    torch.cuda.reset_peak_memory_stats()
    for duration in time_augmentor(aug, batches, args.repeat, torch.cuda.synchronize):
        print("\nDURATION: {} sec\n".format(duration))
        print("PEAK DEVICE MEMORY: {:.1f} MiB".format(torch.cuda.max_memory_allocated() / 2 ** 20))
    print(f"last batch: {int(batches[-1][1].shape[1])} directed edges in, num_remove={aug.num_remove}")


if __name__ == "__main__":
    main()
