#!/usr/bin/env python
"""bench.py — rLap views/s on the shapes BASELINE.json names (default: the ogbn-arxiv-shaped headline, configs[3]).

  python bench.py [--config C1|C2|C3|C4|C5] [--o_v ..] [--o_n ..] [--gpus N] [--steps K] [--warmup W] [--views V]
                  [--impl ours|reference]

One "step" = one pass of the hot path over one batch: ingest (COO -> CSR, validation) of the config's synthetic
graph plus V independent views (ordering, elimination, emission) on every GPU. Views are sharded over GPUs by view id
(weak scaling: V views per GPU per step, no data-path collective). `value` is whole-job views/s with the edge list
resident in HBM; `e2e` is the same through the public API with HOST buffers (pinned edge_index in, packed rows out,
copies inside the timed region). `--impl reference` times the reference's own CPU implementation (oracle/_ref, the
unmodified C++ built against a container-only Eigen stand-in) on all host cores.

Configs (BASELINE.json `configs`):
  C1  BA n=100, m=50, num_remove=50, random/asc (the reference's README / test example), 256 views per step
  C2  Cora-shaped SBM n=2708, 10556 directed edges, num_remove=812 (30 %), any o_v/o_n (default degree/asc), 256 views
  C3  PROTEINS-shaped batch of 1113 graphs (~39 nodes each), 16 views of every graph per step, num_remove = n_g // 2 per
      graph; a "view" is one augmented graph (17808 per step). The reference arm follows the reference's own call
      pattern: unions of 128 graphs per call (scripts/graph_shared.py:139-146, DataLoader(batch_size=128))
  C4  arxiv-shaped BA n=169343, m=7 (~2.37 M directed edges), num_remove=50 %, degree/asc, 128 views per step (headline;
      round 1 ran 64 per step: `--views 64` reproduces that line)
  C5  products-shaped SBM n=2449029, 123.7 M directed edges, coarsen, num_remove=50 %, 4 views per step
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rlap_views_per_sec"
UNIT = "views/s"

CONFIGS = {
    "C1": dict(o_v="random", o_n="asc", views=256,
               workload="C1 Barabasi-Albert graph n=100 m=50 (README/tests example), num_remove=50"),
    "C2": dict(o_v="degree", o_n="asc", views=256,
               workload="C2 Cora-shaped SBM n=2708 E=10556 directed, num_remove=812 (30%)"),
    "C3": dict(o_v="random", o_n="asc", views=16,
               workload="C3 PROTEINS-shaped batch of 1113 graphs (~39 nodes each), num_remove=50% per graph, one view = one augmented graph"),
    "C4": dict(o_v="degree", o_n="asc", views=128,
               workload="C4 arxiv-shaped BA graph n=169343 E~2.37M directed, num_remove=50%"),
    "C5": dict(o_v="coarsen", o_n="asc", views=4,
               workload="C5 products-shaped SBM n=2449029 E~123.7M directed, num_remove=50%"),
}


class Workload:
    """the synthetic input of one config: edge_index (numpy on the host, or a torch tensor for C5), node count, graph
    pointer of a batch, removals, and how many "views" one view of the whole input counts for"""

    def __init__(self, name, device=None):
        from rlap_b200 import graphs
        self.name = name
        self.graph_ptr = None
        self.units = 1
        self.ei_torch = None
        if name == "C1":
            self.n, self.ei = 100, graphs.barabasi_albert(100, 50, seed=0)
            self.num_remove = 50
        elif name == "C2":
            self.n, self.ei = 2708, graphs.sbm(2708, 7, 5278, seed=0)
            self.num_remove = 812
        elif name == "C3":
            self.ei, ptr = graphs.proteins_like_batch(1113, seed=0)
            self.n, self.graph_ptr = int(ptr[-1]), ptr
            self.num_remove = np.diff(ptr) // 2
            self.units = 1113
        elif name == "C4":
            self.n = 169343
            self.ei = _cached(f"/tmp/rlap_b200_ba_{self.n}_7_seed0.npy", lambda: graphs.barabasi_albert(self.n, 7, seed=0))
            self.num_remove = self.n // 2
        elif name == "C5":
            import torch
            self.n = 2449029
            dev = device if device is not None else "cpu"
            self.ei_torch = graphs.sbm_torch(self.n, 47, 61859140, seed=0, device=dev)
            self.ei = None
            self.num_remove = self.n // 2
        else:
            raise ValueError(name)
        self.E = int(self.ei.shape[1]) if self.ei is not None else int(self.ei_torch.shape[1])

    def edge_index_numpy(self):
        return self.ei if self.ei is not None else self.ei_torch.cpu().numpy()


def _cached(path, make):
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    arr = make()
    try:
        np.save(path + f".{os.getpid()}.tmp.npy", arr)
        os.replace(path + f".{os.getpid()}.tmp.npy", path)
    except Exception:
        pass
    return arr


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons with NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as ex:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference C++ on the host cores
# ----------------------------------------------------------------------------------------------
def _ref_calls(wl):
    """the reference's call pattern for one view of the whole input: [(edge_info, n, t)] (C3: unions of 128 graphs)"""
    ei = wl.edge_index_numpy()
    if wl.graph_ptr is None:
        info = np.concatenate([ei.T.astype(np.float64), np.ones((ei.shape[1], 1))], axis=1)
        return [(info, wl.n, int(wl.num_remove))]
    calls = []
    ptr = wl.graph_ptr
    G = len(ptr) - 1
    owner = np.searchsorted(ptr, ei[1], side="right") - 1
    for g0 in range(0, G, 128):
        g1 = min(g0 + 128, G)
        lo, hi = int(ptr[g0]), int(ptr[g1])
        sel = (owner >= g0) & (owner < g1)
        sub = ei[:, sel] - lo
        info = np.concatenate([sub.T.astype(np.float64), np.ones((sub.shape[1], 1))], axis=1)
        calls.append((info, hi - lo, int(0.5 * (hi - lo))))      # num_remove = int(frac * num_nodes) of the union
    return calls


def _ref_worker(args):
    path, o_v, o_n, views = args
    from oracle import ref
    data = np.load(path, allow_pickle=True)
    calls = [(np.ascontiguousarray(c[0]), int(c[1]), int(c[2])) for c in data]
    t0 = time.perf_counter()
    rows = 0
    for _ in range(views):
        for info, n, t in calls:
            rows += ref.approximate_cholesky(info, n, t, o_v, o_n).shape[0]
    return time.perf_counter() - t0, rows


def cpu_reference_throughput(calls, o_v, o_n, procs, views_per_proc=1):
    """one process per core, each producing `views_per_proc` views of the whole input with the reference build;
    returns (views of the whole input per second, wall seconds)"""
    import multiprocessing as mp
    path = f"/tmp/rlap_b200_calls_{os.getpid()}.npy"
    arr = np.empty(len(calls), dtype=object)
    for i, c in enumerate(calls):
        arr[i] = c
    np.save(path, arr, allow_pickle=True)
    try:
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            res = pool.map(_ref_worker, [(path, o_v, o_n, views_per_proc)] * procs, chunksize=1)
        # all workers run concurrently; the slowest one bounds the batch (process start-up and the load of
        # the input file are not charged to the reference)
        wall = max(r[0] for r in res)
    finally:
        os.remove(path)
    return procs * views_per_proc / wall, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config_dict(wl, o_v, o_n, **extra):
    d = {"workload": f"{CONFIGS[wl.name]['workload']}, o_v={o_v}, o_n={o_n}", "config": wl.name, "edges": wl.E,
         "nodes": wl.n}
    d.update(extra)
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    wl = Workload(args.config)
    o_v, o_n = args.o_v, args.o_n
    cores = min(host_cores(), 64)
    if args.config == "C5":
        cores = min(cores, 2)          # 3 GB of float64 input + one heap node per directed edge per process
    calls = _ref_calls(wl)
    single = None
    if not ref.available():
        # oracle port (keyed mode, single thread) stands in when the reference build is absent
        from oracle import port
        ei = wl.edge_index_numpy()
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            ptr, col, w = port.ingest(ei, None, wl.n)
            port.keyed_schur(ptr, col, w, wl.num_remove, o_v, o_n, seed=s, graph_ptr=wl.graph_ptr)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        val, kind, cores = wl.units / statistics.mean(times), "port", 1
        sample = "1 view of the input per step, oracle keyed mode"
        ms = 1e3 * statistics.mean(times)
    else:
        # the parent loads the library too (the forked workers inherit the mapping): the single-process figure is timed here
        ref.lib()
        per = max(1, int(round(0.05 / max(_time_one(calls, o_v, o_n), 1e-6))))   # ~50 ms of work per worker and step at least
        per = min(per, 2000)
        if wl.name in ("C4", "C5"):
            per = 1
        sp = []
        for s in range(4 if wl.name not in ("C4", "C5") else 2):
            t0 = time.perf_counter()
            for info, n, t in calls:
                ref.approximate_cholesky(info, n, t, o_v, o_n)
            sp.append(time.perf_counter() - t0)
        single = wl.units / statistics.median(sp[1:])
        times = []
        for s in range(args.warmup + args.steps):
            v, wall = cpu_reference_throughput(calls, o_v, o_n, cores, per)
            if s >= args.warmup:
                times.append(wall)
        ms = 1e3 * statistics.mean(times)
        val, kind = wl.units * cores * per / statistics.mean(times), "reference"
        sample = (f"{cores} processes x {per} view(s) of the input per step (unmodified reference C++, Eigen replaced by a "
                  f"container stand-in); {len(calls)} call(s) per view")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl, o_v, o_n),
        "edges_per_sec": val / wl.units * wl.E,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "single_process_value": single},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _time_one(calls, o_v, o_n):
    from oracle import ref
    t0 = time.perf_counter()
    for info, n, t in calls:
        ref.approximate_cholesky(info, n, t, o_v, o_n)
    return time.perf_counter() - t0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
# kernel launches of one step (counted from the launch code): ingest 12 (count, scatter, rows_warp, rows_block, compact,
# symmetry, 2 x 3 scan kernels) + export; views: setup_graphs, k_eliminate (ONE cooperative launch for all view groups),
# combine_groups, emission count pass (prep, scatter, fsort, base, sort_small, 4 x sort_mid, sort_block, sort_big, 2 x 3
# scan kernels), export_views, k_emit_write
LAUNCHES_PER_STEP = 13 + 3 + 17 + 1 + 1


def run_ours(args):
    import torch
    import torch.distributed as dist
    import rlap_b200
    from rlap_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; rlap_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args.config, device=dev)
    o_v, o_n = args.o_v, args.o_n
    E, n, V = wl.E, wl.n, args.views
    t = wl.num_remove
    if wl.ei_torch is not None:
        ei_dev = wl.ei_torch
        ei_pinned = torch.empty(ei_dev.shape, dtype=torch.int64).pin_memory()
        ei_pinned.copy_(ei_dev)
    else:
        ei_pinned = torch.from_numpy(wl.ei).pin_memory()
        ei_dev = ei_pinned.to(dev)
    launches = {"n": 0}
    stats_acc = []

    def step_device(step):
        g = ops.prepare(ei_dev, None, n, graph_ptr=wl.graph_ptr)
        out, vp, st = ops.schur_views(g, t, o_v, o_n, num_views=V, seed=1234 + step, view_base=rank * V, dtype=None,
                                      return_stats=True)
        stats_acc.append(st)
        launches["n"] += LAUNCHES_PER_STEP
        return out, vp

    # e2e: what a training loop that prefetches views does. Pinned host edge_index in, packed rows out to pinned
    # host buffers; the device->host copy of step i runs on a copy stream while step i+1 computes (two buffer
    # sets). Every step's H2D and D2H are inside the timed region; at the end the host holds the result of every view.
    #   rows:  (row, col, w) int32/int32/f32 of every view                             12 B per row
    #   csc:   (row int32, column pointers int32 [V, n + 1]): the unweighted CSC adjacency of every view, what the
    #          reference's GCL adapters keep (they drop the weights, scripts/augmentor_benchmarks.py:88-96)   4 B per row
    from concurrent.futures import ThreadPoolExecutor
    copy_stream = torch.cuda.Stream(device=dev)
    host_bufs = [dict(), dict()]
    pending = [None, None]
    pool = ThreadPoolExecutor(max_workers=2)
    d2h_bytes = {"n": 0}

    def finish(done, total, keepalive):
        # the device tensors stay referenced until their copies are done (no record_stream: the caching allocator
        # would hold their blocks back behind cross-stream events and fall back to cudaMalloc, which synchronises)
        done.synchronize()
        del keepalive
        return total

    def make_e2e(mode):
        def step_e2e(step):
            slot = step & 1
            if pending[slot] is not None:          # the buffers of this slot are still being filled by step - 2
                pending[slot].result()
                pending[slot] = None
            d = ei_pinned.to(dev, non_blocking=True)
            g = ops.prepare(d, None, n, graph_ptr=wl.graph_ptr)
            if mode == "rows":
                (row, col, w), vp = ops.schur_views(g, t, o_v, o_n, num_views=V, seed=1234 + step, view_base=rank * V, dtype=None)
                parts = {"row": row, "col": col, "w": w}
            else:
                (row, cp, w), vp = ops.schur_views(g, t, o_v, o_n, num_views=V, seed=1234 + step, view_base=rank * V,
                                                   dtype=None, colptr=True, weights=False)
                parts = {"row": row, "colptr": cp}
            total = int(vp[-1])
            hb = host_bufs[slot]
            nbytes = 0
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                for k, src in parts.items():
                    flat = src.reshape(-1)
                    if k not in hb or hb[k].numel() < flat.numel() or hb[k].dtype != flat.dtype:
                        hb[k] = torch.empty(int(flat.numel() * 1.05) + 16, dtype=flat.dtype).pin_memory()
                    hb[k][:flat.numel()].copy_(flat, non_blocking=True)
                    nbytes += flat.numel() * flat.element_size()
                done = torch.cuda.Event()
                done.record()
            d2h_bytes[mode] = nbytes
            pending[slot] = pool.submit(finish, done, total, (parts, g, d))
            return total
        return step_e2e

    def drain_e2e():
        for slot in (0, 1):
            if pending[slot] is not None:
                pending[slot].result()
                pending[slot] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, e2e=False):
        res = None
        for s in range(warmup):
            res = fn(s)       # same allocation pattern as the timed loop: the previous result stays alive during a step
        if e2e:
            drain_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            res = fn(warmup + s)
        if e2e:
            drain_e2e()          # the last copies are part of the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, res

    sampler = ClockSampler(local)
    sampler.start()
    stats_acc.clear()
    launches["n"] = 0
    ms_total, _ = timed(step_device, args.steps, args.warmup)
    timed_stats = stats_acc[args.warmup:]
    n_launch = launches["n"] * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop()
    if args.skip_e2e:
        ms_e2e = ms_csc = float("nan")
    else:
        ms_e2e, _ = timed(make_e2e("rows"), args.steps, max(args.warmup, 1), e2e=True)
        host_bufs[0].clear(); host_bufs[1].clear()
        ms_csc, _ = timed(make_e2e("csc"), args.steps, max(args.warmup, 1), e2e=True)
        host_bufs[0].clear(); host_bufs[1].clear()

    # single-call latency of the reference's own call shape (one view through ops.approximate_cholesky, device
    # resident edge_index in, [E',3] float64 out), the figure to put beside the reference's per-call time
    lat = []
    if wl.graph_ptr is None and wl.name != "C5":
        nr = int(t)
        for s in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ops.approximate_cholesky(ei_dev, None, n, nr, o_v, o_n, seed=s)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = lat[3:]

    units = wl.units
    ms_step = ms_total / args.steps
    value = world * V * units * args.steps / (ms_total / 1e3)
    e2e_value = None if args.skip_e2e else world * V * units * args.steps / (ms_e2e / 1e3)
    csc_value = None if args.skip_e2e else world * V * units * args.steps / (ms_csc / 1e3)

    # roofline of the dominant kernel (k_eliminate: ordering + elimination of all views of a step),
    # algorithmic bytes per SURVEY.md §8(d): 8n (ordering) + 8D (adjacency read once) + 24F (fill edges, both
    # directions) per view, with D and F counted by the kernel itself
    peak, peak_src = peaks()
    elim_us = statistics.mean(s["elim_us"] for s in timed_stats)
    count_us = statistics.mean(s["emit_count_us"] for s in timed_stats)
    D = statistics.mean(s["raw_entries_read"] for s in timed_stats)
    F = statistics.mean(s["fills"] for s in timed_stats)
    rows = statistics.mean(s["rows"] for s in timed_stats)
    alg_bytes = 8.0 * n * V + 8.0 * D + 24.0 * F
    achieved = alg_bytes / (elim_us * 1e-6) / 1e9
    # whole path, per step: ingest 20E + 4n, then per view 8n + 8D + 24F + 12M + 12E' (M ~ E')
    path_bytes = 20.0 * E + 4.0 * n + alg_bytes + 24.0 * rows
    path_gbs = path_bytes / (ms_step * 1e-3) / 1e9

    # measured DRAM traffic of the same kernel from the committed ncu --set full capture (same workload and V)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_k_eliminate_traffic.json")))
        if int(tj.get("views_per_launch", 0)) == V and tj.get("config") == wl.name and tj.get("o_v") == o_v:
            traffic = float(tj["traffic_bytes_per_launch"])
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(wl, o_v, o_n, views_per_gpu_per_step=V, parallelism=f"views sharded over {world} GPU(s)",
                              step="ingest + V views (ordering, elimination, emission), packed int32/int32/f32 rows",
                              l2="per-step working set exceeds the 126 MB L2" if V * (E * 36 + n * 60) > 126e6 else
                                 "per-step working set fits the L2; every step writes fresh buffers (seed changes per step)"),
        "edges_per_sec": value / units * E,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ei_pinned.numel() * 8),
                "d2h_bytes_per_step": int(d2h_bytes.get("rows", 0)), "ms_per_step": None if args.skip_e2e else ms_e2e / args.steps, "mode": "rows",
                "api": "ops.prepare + ops.schur_views from pinned host edge_index; packed (row, col, w) rows copied back to "
                       "pinned host on a copy stream, overlapping the next step (2 buffer sets)"},
        "e2e_csc_unweighted": {"value": csc_value, "unit": UNIT, "d2h_bytes_per_step": int(d2h_bytes.get("csc", 0)),
                               "ms_per_step": None if args.skip_e2e else ms_csc / args.steps,
                               "api": "the same with schur_views(colptr=True, weights=False): the host receives the unweighted "
                                      "CSC adjacency (row ids + column pointers) of every view, the form the reference's GCL "
                                      "adapters keep (they drop the weights, scripts/augmentor_benchmarks.py:88-96)"},
        "single_call_latency_ms": (statistics.median(lat) if lat else None),
        "gpu_launches": int(n_launch),
        "roofline": {"bound": "hbm", "kernel": "k_eliminate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": elim_us / 1e3,
                     "launches_per_step": 1,
                     "note": "k_eliminate is ONE cooperative launch per step whose blocks are partitioned into view groups "
                             "(own barrier each); kernel_ms is its CUDA-event time and the bytes are those of all views of "
                             "the step",
                     "kernel_share_of_step": elim_us / 1e3 / ms_step,
                     "emit_count_ms": count_us / 1e3,
                     "path": {"algorithmic_bytes_per_step": path_bytes, "achieved": path_gbs, "frac": path_gbs / peak}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path on this box's host cores, in a clean subprocess (bounded sample)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                                "--o_v", o_v, "--o_n", o_n, "--steps", "3", "--warmup", "1"],
                               capture_output=True, text=True, timeout=1500)
            ref_line = json.loads(r.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref_line["cpu_baseline"]
        except Exception as ex:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--o_v", default=None, choices=["random", "degree", "coarsen"])
    ap.add_argument("--o_n", default=None, choices=["asc", "desc", "random"])
    ap.add_argument("--views", type=int, default=None, help="views per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="auxiliary runs only: no host-buffer loop (e2e is reported as null)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.o_v = args.o_v or cfg["o_v"]
    args.o_n = args.o_n or cfg["o_n"]
    args.views = args.views or cfg["views"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
