#!/usr/bin/env python
"""bench.py -- paths tracked / second of the batched path tracker (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: every path of the workload is tracked from
t=1 to t=0 and classified.  `value` = whole-job paths/s with inputs already resident in HBM
(kernel time, CUDA events on the launching stream); `e2e` = the same through hc_track_batch with
HOST buffers (H2D of starts/parameters and D2H of the PathResult SoA inside the timed region).
`--impl reference` times the reference algorithm's CPU restatement (oracle/) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


WORKLOADS = ("cyclic7_polyhedral", "katsura8", "cyclic7_td", "tritangents", "cyclooctane_td", "biochem_sweep")


def make_workload(name, replicas, api=None):
    """The BASELINE.json configs (hcb200.workloads); `replicas` multiplies the small ones so that one
    B200 has enough independent paths (identical work distribution per replica)."""
    import hcb200
    from hcb200 import workloads
    if name == "cyclic7_polyhedral":      # configs[1], the default: 924 mixed-volume paths per replica
        return workloads.cyclic_polyhedral(7, replicas)
    if name == "katsura8":                # configs[0]
        return workloads.katsura8(replicas)
    if name == "cyclic7_td":
        return workloads.cyclic7_total_degree(replicas)
    if name == "tritangents":             # configs[2]: 110 592 paths
        return workloads.tritangents_total_degree()
    if name == "cyclooctane_td":          # configs[3] on the total-degree start system: 32 768 paths
        return workloads.cyclooctane_total_degree()
    if name == "biochem_sweep":           # configs[4]: `replicas` x 1024 parameter points per GPU
        return workloads.biochem_sweep(api, replicas * 1024)
    raise SystemExit(f"unknown workload {name}; choose from {WORKLOADS}")


def cpu_baseline(w, budget_paths, threads, fast=True):
    """The oracle (kind 'port') on the host cores, on a bounded sample of the same workload."""
    import pyoracle
    api = pyoracle.load(fast=fast)
    sample = w.subset(budget_paths)
    handles = sample.build(api)
    t0 = time.perf_counter()
    r = sample.track(api, handles, nthreads=threads)
    dt = time.perf_counter() - t0
    return sample.N / dt, sample, r, dt


def run_reference(args, rank, world):
    import hcb200  # noqa: F401
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import pyoracle
    w = make_workload(args.workload, args.replicas, pyoracle.load(fast=True))
    per_step = args.cpu_sample
    for _ in range(args.warmup):
        cpu_baseline(w, max(64, per_step // 8), threads)
    times, n = [], 0
    for _ in range(args.steps):
        v, sample, r, dt = cpu_baseline(w, per_step, threads)
        times.append(dt); n = sample.N
    value = n * len(times) / sum(times)
    line = {"impl": "reference", "metric": "paths tracked/sec", "value": value, "unit": "paths/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.description, "sample": f"first {n} paths per step"},
            "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "port",
                             "sample": f"first {n} paths of the workload per step, {threads} std::threads, -O3 oracle (reference cannot run: no Julia)"},
            "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cyclic7_polyhedral", choices=WORKLOADS)
    ap.add_argument("--replicas", type=int, default=480, help="replicas of the config per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=131072, help="paths per step of the CPU baseline / reference arm (about 5-10 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import hcb200
    from hcb200 import capi, flops, lib
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's version banner / debug log goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    api = lib.load(local)
    raw = api.raw
    # the job = `replicas x world` replicas of the config; this rank tracks its contiguous shard of
    # the path index range (hcb200.sharding), no collective until the final reduction of the counts
    from hcb200 import sharding
    wg = make_workload(args.workload, args.replicas * (world if args.workload not in ("tritangents", "cyclooctane_td") else 1), api)
    lo, hi = sharding.shard_range(wg.N, rank, world)
    w = wg.slice(lo, hi)
    w.expected = wg.expected
    handles = w.build(api)
    opts = api.default_options()
    dp = lambda a: a.ctypes.data_as(capi.c_double_p)

    # ---- device-resident arm (value): upload once, run K times
    starts = np.ascontiguousarray(w.starts)
    t1 = np.array([1.0, 0.0]); t0 = np.array([0.0, 0.0])
    pq = np.ascontiguousarray(w.path_q).view(np.float64).reshape(-1) if w.path_q is not None else None
    ci = np.ascontiguousarray(w.cell_index, dtype=np.int32) if w.cell_index is not None else None
    cw = np.ascontiguousarray(w.cell_weights, dtype=np.float64) if w.cell_weights is not None else None
    res_h = raw.hc_resident_create(handles["H"].handle, handles["Hcoeff"].handle if "Hcoeff" in handles else None, C.byref(opts), w.mode,
                                   w.N, dp(starts.view(np.float64)), dp(t1), dp(t0), None, dp(pq) if pq is not None else None,
                                   ci.ctypes.data_as(capi.c_int32_p) if ci is not None else None, dp(cw) if cw is not None else None,
                                   cw.shape[0] if cw is not None else 0)
    if not res_h:
        raise SystemExit("hc_resident_create failed: " + raw.hc_last_error().decode())
    res_h = C.c_void_p(res_h)
    ms = C.c_double()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        assert raw.hc_resident_run(res_h, C.byref(ms)) == 0, raw.hc_last_error()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms = []
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        assert raw.hc_resident_run(res_h, C.byref(ms)) == 0, raw.hc_last_error()
        kernel_ms.append(ms.value)
    barrier()
    wall = time.perf_counter() - tw0
    clocks = sampler.stop() if rank == 0 else None
    dev_s = sum(kernel_ms) * 1e-3
    res = capi.BatchResults.allocate(w.n, w.N)
    d = res.desc()
    assert raw.hc_resident_fetch(res_h, C.byref(d)) == 0
    raw.hc_resident_destroy(res_h)
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s_max, wall_max = float(t[0]), float(t[1])
    total_paths = wg.N * args.steps
    value = total_paths / dev_s_max

    # ---- end-to-end arm: host buffers through the public call.  Inputs and the result arrays are allocated and
    # page-locked once (hc_host_register), as a host that solves repeatedly keeps them; every step still copies
    # all inputs host -> device and all result arrays device -> host inside hc_track_batch.
    w.starts = np.ascontiguousarray(w.starts, dtype=np.complex128)
    if w.path_q is not None:
        w.path_q = np.ascontiguousarray(w.path_q, dtype=np.complex128)
    if w.cell_index is not None:
        w.cell_index = np.ascontiguousarray(w.cell_index, dtype=np.int32)
    out = capi.BatchResults.allocate(w.n, w.N)
    if w.sweep_starts is not None:   # many_solve entry point: k starts + one parameter column per point cross the bus
        inputs = [w.sweep_starts, w.sweep_q]
    else:
        inputs = [w.starts, w.path_q, w.cell_index]
    pinned, pin_note = [], "pageable"
    if not os.environ.get("HC_BENCH_PAGEABLE"):
        try:
            pinned = lib.pin(*inputs, *out.arrays())
            pin_note = "page-locked once (hc_host_register), reused every step"
        except RuntimeError as e:   # e.g. a locked-memory limit on the box: measure with pageable buffers and say so
            pin_note = f"pageable ({e})"
    for _ in range(2):
        w.track(api, handles, opts, out=out)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        r = w.track(api, handles, opts, out=out)
    barrier()
    te = torch.tensor([time.perf_counter() - te0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    tm = lib.timing()
    e2e = total_paths / float(te[0])
    e2e_same = bool((r.return_code == res.return_code).all() and np.array_equal(r.solution, res.solution))
    lib.unpin(pinned)

    # ---- final reduction of the solution-class counts (the only cross-rank exchange of the job)
    counts = sharding.class_counts(res)
    ct = torch.tensor([counts[k] for k in sorted(counts)], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(ct)
    counts = {k: int(v) for k, v in zip(sorted(counts), ct.tolist())}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant (only) kernel
    n_ok = counts["success"]
    fl = flops.batch_flops(w.costs, res.counters, res.accepted_steps, res.rejected_steps)
    peak_gflops = raw.hc_dfma_peak(200000)
    ach_tflops = fl / (np.mean(kernel_ms) * 1e-3) / 1e12
    hbm_peak, hbm_src = load_peaks()
    alg_bytes = flops.path_bytes(w.n) * w.N
    traffic, traffic_src = None, None
    try:  # measured DRAM bytes per path of this workload's kernel (one ncu --set full capture, profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tj:
            traffic, traffic_src = tj["dram_bytes_per_path"] * w.N, tj["capture"]
    except (OSError, ValueError, KeyError):
        pass
    line = {
        "metric": "paths tracked/sec", "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wg.description, "paths_per_gpu_per_step": w.N,
                   "parallelism": f"path index range sharded over {world} GPU(s), no collective on the hot path",
                   "l2": (f"per-lane path state in local memory, {tm.slab_bytes / 2**20:.0f} MiB over all lanes > 126 MiB L2; inputs are KBs" if tm.lanes == 1 else
                          f"path state in shared memory ({tm.slab_bytes} B per path); inputs are KBs"),
                   "engine": "thread-per-path" if tm.lanes == 1 else f"{tm.lanes}-lane group per path",
                   "grid": tm.grid, "block": tm.block, "success_paths": n_ok, "class_counts": counts, "expected": wg.expected},
        "e2e": {"value": e2e, "unit": "paths/s", "h2d_bytes_per_step": int(tm.h2d_bytes), "d2h_bytes_per_step": int(tm.d2h_bytes),
                "last_call_ms": {"setup_and_h2d": round(tm.h2d_ms, 2), "kernel": round(tm.kernel_ms, 2), "d2h": round(tm.d2h_ms, 2),
                                 "whole_step_mean": round(1e3 * float(te[0]) / args.steps, 2)},
                "entry_point": "hc_track_sweep" if w.sweep_starts is not None else ("hc_polyhedral_track_batch" if w.mode == 2 else "hc_track_batch"),
                "host_buffers": pin_note,
                "results_identical_to_resident_arm": e2e_same},
        "gpu_launches": args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": ach_tflops, "peak": peak_gflops / 1e3, "unit": "TFLOP/s",
                     "frac": ach_tflops / (peak_gflops / 1e3), "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "hc_dfma_peak microbenchmark measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                     "flops_per_path": fl / w.N,
                     "hbm": {"algorithmic_gbs": alg_bytes / (np.mean(kernel_ms) * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src}},
        "wall_s_timed_region": wall_max,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, sample, rc, dt = cpu_baseline(w, args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": v, "unit": "paths/s", "cores": threads, "kind": "port",
                                "sample": f"first {sample.N} paths of the workload, {threads} std::threads, -O3 oracle, {dt:.1f} s"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
