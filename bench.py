#!/usr/bin/env python
"""bench.py — rLap views/s on the ogbn-arxiv-shaped graph (BASELINE.json configs[3]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--views V] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: ingest (COO -> CSR, validation) of the
arxiv-shaped Barabasi-Albert graph (169,343 nodes, ~2.37M directed edges, synthetic, seeded) plus
V independent degree/asc views with num_remove = 50 % (ordering, elimination, emission), on every GPU.
Views are sharded over GPUs by view id (weak scaling: V views per GPU per step, no data-path
collective). `value` is whole-job views/s with the edge list resident in HBM; `e2e` is the same
through the public API with HOST buffers (pinned edge_index in, packed rows out, copies inside the
timed region). `--impl reference` times the reference's own CPU implementation (oracle/_ref, the
unmodified C++ built against a container-only Eigen stand-in) on all host cores.
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

N_NODES = 169343
BA_M = 7
O_V, O_N = "degree", "asc"
METRIC = "rlap_views_per_sec"
UNIT = "views/s"


def make_graph():
    """arxiv-shaped BA graph, cached under /tmp (generation is a python loop of a few seconds)"""
    path = f"/tmp/rlap_b200_ba_{N_NODES}_{BA_M}_seed0.npy"
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    from rlap_b200 import graphs
    ei = graphs.barabasi_albert(N_NODES, BA_M, seed=0)
    try:
        np.save(path + f".{os.getpid()}.tmp.npy", ei)
        os.replace(path + f".{os.getpid()}.tmp.npy", path)
    except Exception:
        pass
    return ei


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
def _ref_worker(args):
    info_path, n, t, views = args
    from oracle import ref
    info = np.load(info_path, mmap_mode="r")
    info = np.ascontiguousarray(info)
    t0 = time.perf_counter()
    rows = 0
    for _ in range(views):
        rows += ref.approximate_cholesky(info, n, t, O_V, O_N).shape[0]
    return time.perf_counter() - t0, rows


def cpu_reference_throughput(ei, procs, views_per_proc=1):
    """one process per core, each producing `views_per_proc` views with the reference build;
    returns (views/s, wall seconds, kind)"""
    import multiprocessing as mp
    from oracle import ref
    if not ref.available():
        return None
    info = np.concatenate([ei.T.astype(np.float64), np.ones((ei.shape[1], 1))], axis=1)
    path = f"/tmp/rlap_b200_info_{os.getpid()}.npy"
    np.save(path, info)
    try:
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            res = pool.map(_ref_worker, [(path, N_NODES, N_NODES // 2, views_per_proc)] * procs, chunksize=1)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    ei = make_graph()
    cores = min(host_cores(), 64)
    if not ref.available():
        # oracle port (keyed mode, single thread) stands in when the reference build is absent
        from oracle import port
        ptr, col, w = port.ingest(ei, None, N_NODES)
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            port.ingest(ei, None, N_NODES)
            port.keyed_schur(ptr, col, w, N_NODES // 2, O_V, O_N, seed=s)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        val, kind, cores, sample = 1.0 / statistics.mean(times), "port", 1, "1 view per step, oracle keyed mode"
        ms = 1e3 * statistics.mean(times)
    else:
        times = []
        for s in range(args.warmup + args.steps):
            v, wall = cpu_reference_throughput(ei, cores, 1)
            if s >= args.warmup:
                times.append(wall)
        ms = 1e3 * statistics.mean(times)
        val, kind = cores / statistics.mean(times), "reference"
        sample = f"{cores} processes x 1 view per step (unmodified reference C++, Eigen replaced by a container stand-in)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 arxiv-shaped BA graph n=169343 E~2.37M directed, num_remove=50%, o_v=degree, o_n=asc",
                   "edges": int(ei.shape[1])},
        "edges_per_sec": val * int(ei.shape[1]),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
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

    ei_np = make_graph()
    E = int(ei_np.shape[1])
    V = args.views
    t = N_NODES // 2
    ei_pinned = torch.from_numpy(ei_np).pin_memory()
    ei_dev = ei_pinned.to(dev)
    launches = {"n": 0}
    stats_acc = []

    step_wall = []

    def step_device(step):
        step_wall.append(time.perf_counter())
        g = ops.prepare(ei_dev, None, N_NODES)
        out, vp, st = ops.schur_views(g, t, O_V, O_N, num_views=V, seed=1234 + step, view_base=rank * V, dtype=None,
                                      return_stats=True)
        stats_acc.append(st)
        # ingest: 12 kernel launches; views: setup, k_eliminate once per view group (min(V, 32) concurrent cooperative
        # launches) + combine, prep, base, scatter, sort_warp, 4 x sort_mid, sort_block, sort_big, 2 x 3 scan, export, copy
        groups = V // ((V + 31) // 32)
        launches["n"] += 12 + 19 + groups + (1 if groups > 1 else 0)
        return out, vp

    # e2e: what a training loop that prefetches views does. Pinned host edge_index in, packed rows out to pinned
    # host buffers; the device->host copy of step i runs on a copy stream while step i+1 computes (two buffer
    # sets). Every step's H2D and D2H are inside the timed region; at the end the host holds (row, col, w) of every
    # view. The link is the bottleneck (26 MB per view). Shipping the column pointers instead of `col`
    # (schur_views(colptr=True) + expand_cols) was measured here too: the D2H drops to 18 MB per view, but rebuilding
    # col with 8-14 host threads competes with the DMA for this host's memory bandwidth and the step got slower
    # (32 / 29 ms against 30 ms), so the plain rows are what is timed.
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

    def step_e2e(step):
        slot = step & 1
        if pending[slot] is not None:          # the buffers of this slot are still being filled by step - 2
            pending[slot].result()
            pending[slot] = None
        d = ei_pinned.to(dev, non_blocking=True)
        g = ops.prepare(d, None, N_NODES)
        (row, col, w), vp = ops.schur_views(g, t, O_V, O_N, num_views=V, seed=1234 + step, view_base=rank * V, dtype=None)
        total = int(vp[-1])
        hb = host_bufs[slot]
        if "row" not in hb or hb["row"].numel() < total:
            cap = int(total * 1.05)
            hb["row"] = torch.empty(cap, dtype=torch.int32).pin_memory()
            hb["col"] = torch.empty(cap, dtype=torch.int32).pin_memory()
            hb["w"] = torch.empty(cap, dtype=torch.float32).pin_memory()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            hb["row"][:total].copy_(row, non_blocking=True)
            hb["col"][:total].copy_(col, non_blocking=True)
            hb["w"][:total].copy_(w, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        d2h_bytes["n"] = total * 12
        pending[slot] = pool.submit(finish, done, total, (row, col, w, g, d))
        return total

    # the same with per-view column pointers instead of the col array over the link (18 instead of 26 MB per view) and
    # col rebuilt in the pinned host buffer by host threads; pays off when several GPUs share the host's ingress
    fill_threads = max(2, min(8, host_cores() // max(world, 1)))

    def finish_colptr(done, hb, vp, total, keepalive):
        done.synchronize()
        del keepalive
        ops.expand_cols(hb["colptr"], vp, out=hb["col"], threads=fill_threads)
        return total

    def step_e2e_colptr(step):
        slot = step & 1
        if pending[slot] is not None:
            pending[slot].result()
            pending[slot] = None
        d = ei_pinned.to(dev, non_blocking=True)
        g = ops.prepare(d, None, N_NODES)
        (row, cp, w), vp = ops.schur_views(g, t, O_V, O_N, num_views=V, seed=1234 + step, view_base=rank * V, dtype=None,
                                           colptr=True)
        total = int(vp[-1])
        hb = host_bufs[slot]
        if "row" not in hb or hb["row"].numel() < total or "colptr" not in hb:
            cap = int(total * 1.05)
            hb["row"] = torch.empty(cap, dtype=torch.int32).pin_memory()
            hb["col"] = torch.empty(cap, dtype=torch.int32).pin_memory()
            hb["w"] = torch.empty(cap, dtype=torch.float32).pin_memory()
            hb["colptr"] = torch.empty((V, N_NODES + 1), dtype=torch.int32).pin_memory()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            hb["row"][:total].copy_(row, non_blocking=True)
            hb["w"][:total].copy_(w, non_blocking=True)
            hb["colptr"].copy_(cp, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        d2h_bytes["n"] = total * 8 + cp.numel() * 4
        pending[slot] = pool.submit(finish_colptr, done, hb, vp, total, (row, cp, w, g, d))
        return total

    e2e_mode = args.e2e_mode if args.e2e_mode != "auto" else ("colptr" if world >= 2 else "rows")
    e2e_step = step_e2e_colptr if e2e_mode == "colptr" else step_e2e

    def drain_e2e():
        for slot in (0, 1):
            if pending[slot] is not None:
                pending[slot].result()
                pending[slot] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        res = None
        for s in range(warmup):
            res = fn(s)       # same allocation pattern as the timed loop: the previous result stays alive during a step
        if fn is e2e_step:
            drain_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            res = fn(warmup + s)
        if fn is e2e_step:
            drain_e2e()          # the last copies (and host rebuilds) are part of the timed region
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
    if os.environ.get("BENCH_DEBUG"):
        print("step starts (ms since first):", [round((x - step_wall[0]) * 1e3, 1) for x in step_wall], file=sys.stderr)
        print("reserved MB", torch.cuda.memory_reserved() >> 20, "allocated MB", torch.cuda.memory_allocated() >> 20,
              "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"),
              "segments", torch.cuda.memory_stats().get("segment.all.allocated"), file=sys.stderr)
    n_launch = launches["n"] * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop()
    ms_e2e, total_rows = timed(e2e_step, args.steps, max(args.warmup, 1))

    ms_step = ms_total / args.steps
    value = world * V * args.steps / (ms_total / 1e3)
    e2e_value = world * V * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel (k_eliminate: ordering + elimination of all views of a step),
    # algorithmic bytes per SURVEY.md §8(d): 8n (ordering) + 8D (adjacency read once) + 24F (fill edges, both
    # directions) per view, with D and F counted by the kernel itself
    peak, peak_src = peaks()
    elim_us = statistics.mean(s["elim_us"] for s in timed_stats)
    count_us = statistics.mean(s["emit_count_us"] for s in timed_stats)
    D = statistics.mean(s["raw_entries_read"] for s in timed_stats)
    F = statistics.mean(s["fills"] for s in timed_stats)
    rows = statistics.mean(s["rows"] for s in timed_stats)
    alg_bytes = 8.0 * N_NODES * V + 8.0 * D + 24.0 * F
    achieved = alg_bytes / (elim_us * 1e-6) / 1e9
    # whole path, per step: ingest 20E + 4n, then per view 8n + 8D + 24F + 12M + 12E' (M ~ E')
    path_bytes = 20.0 * E + 4.0 * N_NODES + alg_bytes + 24.0 * rows
    path_gbs = path_bytes / (ms_step * 1e-3) / 1e9

    # measured DRAM traffic of the same kernel from the committed ncu --set full capture (same workload and V)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01c_k_eliminate_traffic.json")))
        if int(tj.get("views_per_launch", 0)) == V:
            traffic = float(tj["traffic_bytes_per_launch"])
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C4 arxiv-shaped BA graph n=169343 E~2.37M directed, num_remove=50%, o_v=degree, o_n=asc",
                   "edges": E, "views_per_gpu_per_step": V, "parallelism": f"views sharded over {world} GPU(s)",
                   "step": "ingest + V views (ordering, elimination, emission), packed int32/int32/f32 rows",
                   "l2": "per-step working set (~85 MB per view) exceeds the 126 MB L2 for V >= 2"},
        "edges_per_sec": value * E,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ei_pinned.numel() * 8),
                "d2h_bytes_per_step": int(d2h_bytes["n"]), "ms_per_step": ms_e2e / args.steps,
                "mode": e2e_mode,
                "api": ("ops.prepare + ops.schur_views from pinned host edge_index; packed rows copied back to pinned host "
                        "on a copy stream, overlapping the next step (2 buffer sets)") if e2e_mode == "rows" else
                       ("ops.prepare + ops.schur_views(colptr=True) from pinned host edge_index; rows, weights and per-view "
                        "column pointers copied back to pinned host on a copy stream, col rebuilt there by "
                        f"ops.expand_cols with {fill_threads} host threads, both overlapping the next step (2 buffer sets); "
                        "the host ends with (row, col, w) of every view")},
        "gpu_launches": int(n_launch),
        "roofline": {"bound": "hbm", "kernel": "k_eliminate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": elim_us / 1e3,
                     "launches_per_step": V // ((V + 31) // 32),
                     "note": "k_eliminate runs as min(V, 32) concurrent cooperative launches (view groups, one grid "
                             "barrier each); kernel_ms is the CUDA-event time from the first launch to the join of all "
                             "of them and the bytes are those of all views of the step",
                     "kernel_share_of_step": elim_us / 1e3 / ms_step,
                     "emit_count_ms": count_us / 1e3,
                     "path": {"algorithmic_bytes_per_step": path_bytes, "achieved": path_gbs, "frac": path_gbs / peak}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path on this box's host cores, in a clean subprocess (bounded sample: 1 view per core)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                "--warmup", "0"], capture_output=True, text=True, timeout=900)
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
    ap.add_argument("--views", type=int, default=64, help="views per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-mode", default="rows", choices=["rows", "colptr", "auto"],
                    help="what crosses the link in the e2e loop: packed rows, or rows + column pointers (col rebuilt on the host)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
