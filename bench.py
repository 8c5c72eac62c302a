#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend for VGL's frontier-processing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pr|bfs|sssp|cc]

Default workload = BASELINE.json configs[1]: PageRank pull, 20 iterations fp32, RMAT scale-24 edge-factor 16 on one
B200. A "step" is one complete PageRank run (20 sweeps) over the resident graph. Metric = GTEPS in the reference's
convention: iterations x graph edges / time / 1e9 (performance_stats.hpp:272-275, pr.hpp:147).

  value        device-timed (CUDA events on the library's stream), graph resident in HBM, max over ranks
  e2e          the same metric through the C-ABI with HOST buffers: every step copies the VectCSR arrays the reference
               host build owns (row pointers, adjacency, id map; pinned memory) to HBM (vglb_graph_from_csr =
               VGL_Graph::move_to_device), runs vglb_pagerank and copies the rank vector back
  roofline     pr_sweep_kernel: algorithmic bytes per sweep (8E + 16V [+4V on the last]) / average sweep duration,
               against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the reference's own multicore (OpenMP) PageRank (oracle/_ref, unmodified VGL) on this box's host cores,
               on a bounded sample of the same workload

`--impl reference` times the reference's CPU path alone (same metric/config/unit), each step a bounded sample.
Under torchrun (N > 1) every rank owns a 1D vertex range of the graph (multi-GPU).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GTEPS"
WORKLOADS = {
    # name: (generator kind, scale, edge factor, description)   — BASELINE.json configs[1..3]
    "pr": (0, 24, 16, "PageRank pull, 20 iters fp32, RMAT scale-24 ef16"),
    "sssp": (2, 24, 32, "SSSP frontier Bellman-Ford, fp32 weights, uniform-random scale-24 ef32"),
    "bfs": (1, 26, 16, "Direction-optimising BFS, Graph500 Kronecker scale-26 ef16"),
    "cc": (0, 24, 16, "CC min-label hook + jump, RMAT scale-24 ef16 symmetrised"),
}
PR_ITERS = 20


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ncu_traffic(workload: str, world: int):
    """DRAM bytes (read + write) of one launch of the dominant kernel from the committed `ncu --set full` summary of this
    very command (profiles/, B200_PROFILING.md recipe); None where no capture of the configuration exists."""
    if workload != "pr" or world != 1:
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", "r1_pr_sweep_ncu_full.txt")):
            if line.strip().startswith("traffic (dram read + write) bytes:"):
                return int(line.split(":")[1])
    except OSError:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (oracle/_ref) or, where it is not built, the C oracle port
# ---------------------------------------------------------------------------------------------------------------------

def cpu_reference_run(workload: str, sample_scale: int, steps: int, warmup: int):
    """Times the reference's own CPU implementation of `workload` on a bounded sample. Returns (gteps, info dict)."""
    threads = max(2, host_threads())  # the reference segfaults with one OpenMP thread
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    import oracle as O
    kind, _, ef, _ = WORKLOADS[workload]
    V = 1 << sample_scale
    src, dst = O.generate_edges(kind, sample_scale, ef)
    if workload == "cc":
        src, dst = O.symmetrize(src, dst)
    E = len(src)
    outdeg = np.bincount(src, minlength=V)
    sources = O.pick_sources(V, outdeg, max(1, steps + warmup))
    times = []
    if O.ref_available(workload):
        kind_s = "reference"
        rg = O.RefGraph(V, src, dst, workload)
        for i in range(warmup + steps):
            if workload == "pr":
                _, t = rg.pagerank(PR_ITERS)
            elif workload == "bfs":
                _, t = rg.bfs(sources[i], 0)
            elif workload == "sssp":
                _, t = rg.sssp(sources[i], O.MASTER_SEED ^ 0x5555, 2)  # PARTIAL_ACTIVE PUSH (timing only: racy)
            else:
                _, t = rg.cc()
            if i >= warmup:
                times.append(t)
        rg.close()
    else:
        kind_s = "port"
        og = O.OracleGraph(V, src, dst)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            if workload == "pr":
                og.pagerank_f32(PR_ITERS, threads)
            elif workload == "bfs":
                og.bfs(sources[i])
            elif workload == "sssp":
                og.sssp_frontier_bf(sources[i], O.MASTER_SEED ^ 0x5555)
            else:
                og.cc()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    mult = PR_ITERS if workload == "pr" else 1
    t_mean = float(np.mean(times))
    gteps = mult * E / t_mean / 1e9
    gen = {0: "RMAT", 1: "Kronecker", 2: "uniform-random"}[kind]
    info = {"value": gteps, "unit": METRIC, "cores": threads, "kind": kind_s,
            "sample": f"{gen} scale-{sample_scale} ef{ef}" + (" symmetrised" if workload == "cc" else "")
                      + (f", {PR_ITERS} iterations" if workload == "pr" else f", {len(times)} sources" if workload != "cc" else "")
                      + f", {len(times)} timed runs, VGL multicore (OpenMP) build" + ("" if kind_s == "reference" else " restated in C"),
            "seconds_per_run": t_mean}
    return gteps, info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, scale, ef, desc = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    gteps, info = cpu_reference_run(args.workload, args.cpu_scale, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gteps, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": info["seconds_per_run"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.workload in ("pr", "sssp") else "int32", "data": "synthetic",
        "config": {"workload": desc, "sample": info["sample"]},
        "cpu_baseline": info,
        "e2e": {"value": gteps, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

def ours(args):
    import torch
    import vectorgraphlibrary_b200 as vgl
    from vectorgraphlibrary_b200 import dist as vdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libvgl_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    comm = vdist.Communicator.from_env(local_rank) if world > 1 else None

    kind, scale, ef, desc = WORKLOADS[args.workload]
    if args.scale:
        scale = args.scale
    V, E = 1 << scale, ef << scale
    ctx = vgl.Context(local_rank)
    peak, peak_src = measured_peak()
    runner = vdist.make_runner(vgl, ctx, comm, args.workload, kind, scale, ef, PR_ITERS)

    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if comm is not None:
            comm.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        runner.step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stats = []
    ev0.record(stream)
    for i in range(args.steps):
        stats.append(runner.step(args.warmup + i))
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    if comm is not None:
        ms_total = comm.max_float(ms_total)
    ms_per_step = ms_total / args.steps
    edges_per_step = runner.edges_per_step  # whole job, all ranks
    value = edges_per_step / (ms_per_step * 1e-3) / 1e9

    # dominant kernel: average launch duration from the library's own CUDA events around the algorithm loop
    kern_s = float(np.mean([s["seconds"] / max(1, s["dominant_launches"]) for s in stats]))
    kern_bytes = float(np.mean([s["dominant_bytes"] / max(1, s["dominant_launches"]) for s in stats]))
    achieved = kern_bytes / kern_s / 1e9
    launches = int(sum(s["kernel_launches"] for s in stats))

    # end to end through the C ABI with host buffers
    e2e = runner.e2e(max(2, min(args.steps, 5)))
    if comm is not None:
        e2e["seconds"] = comm.max_float(e2e["seconds"])
    e2e_value = edges_per_step / e2e["seconds"] / 1e9

    base_scale = WORKLOADS[args.workload][1]
    workload_name = desc if not args.scale else desc.replace(f"scale-{base_scale}", f"scale-{scale}")
    if world > 1:  # weak scaling: the graph grows with the GPU count, per-GPU work is the single-GPU workload
        workload_name = (workload_name.replace(f"scale-{scale}", f"scale-{runner.scale}")
                         + f" (weak scaling: scale-{scale} per GPU x {world} GPUs)")
    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
            try:
                _, cpu = cpu_reference_run(args.workload, args.cpu_scale, 2, 1)
            except Exception as ex:  # the checker must not take the product's number down with it
                cpu = {"value": None, "unit": METRIC, "cores": host_threads(), "kind": "unavailable", "sample": repr(ex)}
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world > 1 and runner.weak else "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": runner.dtype, "data": "synthetic",
            "config": {"workload": workload_name,
                       "vertices": runner.V_total, "edges": runner.E_total, "iterations_per_step": runner.iters_per_step,
                       "partition": runner.partition, "l2": "inputs larger than L2 (adjacency %.2f GB per GPU, L2 126 MB)" % (runner.adj_bytes_per_gpu / 1e9),
                       "seed": hex(vgl.MASTER_SEED)},
            "roofline": {"bound": "hbm", "kernel": runner.dominant_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload, world) if not args.scale else None, "peak_source": peak_src,
                         "bytes_per_launch": kern_bytes, "ms_per_launch": kern_s * 1e3},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "ms_per_step": e2e["seconds"] * 1e3},
            "gpu_launches": launches, "clocks": clocks,
            # the reference's own "total bandwidth" line: edges x INT_ELEMENTS_PER_EDGE x 4 bytes / time
            # (apps/bfs/bfs.cpp:3 -> 16 B per edge, apps/pr/pr.cpp:3 etc. -> 20 B; performance_stats.hpp:272-275)
            "reference_accounting_gbs": edges_per_step * (16 if args.workload == "bfs" else 20) / (ms_per_step * 1e-3) / 1e9,
            # BFS / SSSP: one seeded source per step; `value` is the harmonic mean over the timed sources (events
            # around the whole loop, host gaps between steps included); min / max come from the library's per-call device timers
            "per_source_gteps": ({"min": edges_per_step / max(s["seconds"] for s in stats) / 1e9,
                                  "max": edges_per_step / min(s["seconds"] for s in stats) / 1e9, "sources": len(stats)}
                                 if args.workload in ("bfs", "sssp") else None),
            "extras": runner.extras(),
        }
        print(json.dumps(line), flush=True)
    runner.close()
    if comm is not None:
        comm.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pr", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=int, default=0, help="override the workload's scale (parity/dev runs only)")
    ap.add_argument("--cpu-scale", type=int, default=22, help="scale of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
