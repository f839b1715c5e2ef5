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


WORKLOADS = ("cyclic7_polyhedral", "katsura8", "cyclic7_td", "tritangents", "cyclooctane_td", "cyclooctane_polyhedral", "biochem_sweep")
CAPTION = "paths tracked/sec"


def make_workload(name, replicas, api=None, points=None):
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
    if name == "cyclooctane_polyhedral":  # configs[3] as the benchmark runs it (solve(F) = polyhedral start system)
        return workloads.cyclooctane_polyhedral()
    if name == "biochem_sweep":           # configs[4]: `points` (default replicas x 1024) parameter points per GPU
        return workloads.biochem_sweep(api, points if points is not None else replicas * 1024)
    raise SystemExit(f"unknown workload {name}; choose from {WORKLOADS}")


def cpu_arm(w, budget_paths, threads):
    """The oracle (kind 'port': C++ restatement of the reference, -O3 -march=native, one std::thread per host core) on
    the first `budget_paths` paths of the workload."""
    import pyoracle
    api = pyoracle.load(fast=True, native=True)
    sample = w.subset(budget_paths) if budget_paths < w.N else w
    handles = sample.build(api)
    t0 = time.perf_counter()
    r = sample.track(api, handles, nthreads=threads)
    dt = time.perf_counter() - t0
    return sample.N / dt, sample, r, dt


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation on all host cores, on the SAME workload as our
    arm (for one GPU: the same paths, every step; for N > 1 GPUs the job is N times larger and each step tracks one
    GPU's share of it, stated in `sample`)."""
    import hcb200  # noqa: F401
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import pyoracle
    api = pyoracle.load(fast=True, native=True)
    w = make_workload(args.workload, args.replicas, api)
    per_step = w.N if args.cpu_sample <= 0 else min(w.N, args.cpu_sample)
    for _ in range(args.warmup):
        cpu_arm(w, max(64, per_step // 16), threads)
    times, n = [], 0
    for _ in range(args.steps):
        v, sample, r, dt = cpu_arm(w, per_step, threads)
        times.append(dt); n = sample.N
    value = n * len(times) / sum(times)
    from hcb200 import result
    line = {"impl": "reference", "metric": CAPTION, "value": value, "unit": "paths/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.description, "paths_per_gpu_per_step": w.N},
            "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "port", "cpu": pyoracle.cpu_model(),
                             "sample": (f"all {n} paths of one GPU's workload per step" if n == w.N else f"first {n} paths per step") +
                                       f", {threads} std::threads, oracle built -O3 -march=native on this host (the Julia reference cannot run: no Julia)",
                             "class_counts": result.statistics(r).asdict()},
            "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class Arm:
    """One workload on this rank's GPU: device-resident timing (value) and the end-to-end call with host buffers."""

    def __init__(self, api, raw, w, opts):
        self.api, self.raw, self.w, self.opts = api, raw, w, opts
        self.handles = w.build(api)

    def resident(self, steps, warmup, barrier=lambda: None):
        from hcb200 import capi, lib
        w, raw = self.w, self.raw
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        starts = np.ascontiguousarray(w.starts)
        t1 = np.array([1.0, 0.0]); t0 = np.array([0.0, 0.0])
        pq = np.ascontiguousarray(w.path_q).view(np.float64).reshape(-1) if w.path_q is not None else None
        ci = np.ascontiguousarray(w.cell_index, dtype=np.int32) if w.cell_index is not None else None
        cw = np.ascontiguousarray(w.cell_weights, dtype=np.float64) if w.cell_weights is not None else None
        h = self.handles
        if w.mode == 2 and w.cells is not None:   # polyhedral start solutions made on the device, as the e2e arm's call does
            i64 = lambda a: a.ctypes.data_as(capi.c_int64_p)
            vol, Hm = np.ascontiguousarray(w.cells["volume"], dtype=np.int64), np.ascontiguousarray(w.cells["H"], dtype=np.int64)
            mu, rr = np.ascontiguousarray(w.cells["mu"], dtype=np.float64), np.ascontiguousarray(w.cells["r"], dtype=np.float64)
            raw.hc_resident_create_cells.restype = C.c_void_p
            raw.hc_resident_create_cells.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(capi.Options), C.c_int64, C.c_int64, C.c_int32,
                                                     capi.c_int64_p, capi.c_int64_p, capi.c_double_p, capi.c_double_p, capi.c_double_p]
            res_h = raw.hc_resident_create_cells(h["H"].handle, h["Hcoeff"].handle, C.byref(self.opts), w.cells_first, w.N, len(vol),
                                                 i64(vol), i64(Hm), dp(mu), dp(rr), dp(cw))
        else:
            res_h = raw.hc_resident_create(h["H"].handle, h["Hcoeff"].handle if "Hcoeff" in h else None, C.byref(self.opts), w.mode,
                                           w.N, dp(starts.view(np.float64)), dp(t1), dp(t0), None, dp(pq) if pq is not None else None,
                                           ci.ctypes.data_as(capi.c_int32_p) if ci is not None else None, dp(cw) if cw is not None else None,
                                           cw.shape[0] if cw is not None else 0)
        if not res_h:
            raise SystemExit("hc_resident_create failed: " + raw.hc_last_error().decode())
        res_h = C.c_void_p(res_h)
        ms = C.c_double()
        for _ in range(warmup):
            assert raw.hc_resident_run(res_h, C.byref(ms)) == 0, raw.hc_last_error()
        barrier()
        kernel_ms = []
        tw0 = time.perf_counter()
        self.launches = 0
        for _ in range(steps):
            assert raw.hc_resident_run(res_h, C.byref(ms)) == 0, raw.hc_last_error()
            kernel_ms.append(ms.value)
            self.launches += 1 + (lib.timing().handoff_paths > 0)   # a two-pass batch launches the lane-group kernel as well
        barrier()
        wall = time.perf_counter() - tw0
        res = capi.BatchResults.allocate(w.n, w.N)
        d = res.desc()
        assert raw.hc_resident_fetch(res_h, C.byref(d)) == 0
        raw.hc_resident_destroy(res_h)
        return kernel_ms, wall, res

    def e2e(self, steps, warmup, barrier=lambda: None):
        """host buffers through the public call.  Inputs and the result arrays are allocated and page-locked once
        (hc_host_register), as a host that solves repeatedly keeps them; every step still copies all inputs host ->
        device and all result arrays device -> host inside the call."""
        from hcb200 import capi, lib
        w = self.w
        w.starts = np.ascontiguousarray(w.starts, dtype=np.complex128)
        if w.path_q is not None:
            w.path_q = np.ascontiguousarray(w.path_q, dtype=np.complex128)
        if w.cell_index is not None:
            w.cell_index = np.ascontiguousarray(w.cell_index, dtype=np.int32)
        out = capi.BatchResults.allocate(w.n, w.N)
        inputs = [w.sweep_starts, w.sweep_q] if w.sweep_starts is not None else [w.starts, w.path_q, w.cell_index]
        pinned, pin_note = [], "pageable"
        if not os.environ.get("HC_BENCH_PAGEABLE"):
            try:
                pinned = lib.pin(*inputs, *out.arrays())
                pin_note = "page-locked once (hc_host_register), reused every step"
            except RuntimeError as e:   # e.g. a locked-memory limit on the box: measure with pageable buffers and say so
                pin_note = f"pageable ({e})"
        for _ in range(warmup):
            w.track(self.api, self.handles, self.opts, out=out)
        barrier()
        te0 = time.perf_counter()
        for _ in range(steps):
            r = w.track(self.api, self.handles, self.opts, out=out)
        barrier()
        dt = time.perf_counter() - te0
        tm = lib.timing()
        lib.unpin(pinned)
        return dt, r, tm, pin_note


def roofline_of(w, res, kernel_ms, peak_gflops, workload_name, engine):
    from hcb200 import flops
    fl = flops.batch_flops(w.costs, res.counters, res.accepted_steps, res.rejected_steps)
    ach = fl / (np.mean(kernel_ms) * 1e-3) / 1e12
    hbm_peak, hbm_src = load_peaks()
    P_pp = 0 if w.path_q is None else w.path_q.shape[1]
    alg_bytes = flops.path_bytes(w.n, P_pp, w.mode == 2) * w.N
    traffic, traffic_src = None, None
    try:  # measured DRAM bytes per path of this workload's kernel (ncu captures, profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload_name)
        if tj and tj.get("engine") == engine:
            traffic, traffic_src = tj["dram_bytes_per_path"] * w.N, tj["capture"]
    except (OSError, ValueError, KeyError):
        pass
    return {"bound": "fp64", "achieved": ach, "peak": peak_gflops / 1e3, "unit": "TFLOP/s", "frac": ach / (peak_gflops / 1e3),
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "hc_dfma_peak microbenchmark measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
            "flops_per_path": fl / w.N,
            "hbm": {"algorithmic_gbs": alg_bytes / (np.mean(kernel_ms) * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src}}


ENGINES = {0: "lane group per path (interpreter)", 1: "thread per path (interpreter)", 2: "thread per path, kernel specialised for the system at run time (NVRTC), lockstep CTAs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cyclic7_polyhedral", choices=WORKLOADS)
    ap.add_argument("--replicas", type=int, default=480, help="replicas of the config per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="paths per step of the CPU arms (default: --impl reference: all paths of one GPU's workload; cpu_baseline: 131072)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config / scaling_c5 blocks (the other BASELINE.json configs)")
    ap.add_argument("--c5-points", type=int, default=1_000_000, help="parameter points of the strong-scaling C5 sweep (whole job)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.cpu_sample is None:
            args.cpu_sample = 0
        run_reference(args, rank, world)
        return
    if args.cpu_sample is None:
        args.cpu_sample = 131072

    import torch
    import hcb200
    from hcb200 import capi, lib, result, sharding
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's version banner / debug log goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    api = lib.load(local)
    raw = api.raw
    opts = api.default_options()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_stats(st):
        keys = sorted(st)
        t = torch.tensor([st[k] for k in keys], dtype=torch.int64, device="cuda")
        if dist is not None:
            dist.all_reduce(t)
        return {k: int(v) for k, v in zip(keys, t.tolist())}

    # the job = `replicas x world` replicas of the config; this rank tracks its contiguous shard of
    # the path index range (hcb200.sharding), no collective until the final reduction of the counts
    fixed = args.workload in ("tritangents", "cyclooctane_td", "cyclooctane_polyhedral")
    wg = make_workload(args.workload, args.replicas * (1 if fixed else world), api)
    lo, hi = sharding.shard_range(wg.N, rank, world)
    w = wg.slice(lo, hi)
    w.expected = wg.expected
    arm = Arm(api, raw, w, opts)
    for _ in range(1):   # the first launch of a kernel on a fresh context pays module load + local-memory setup
        pass
    sampler = ClockSampler(local)
    kernel_ms, wall, res = None, None, None
    # ---- device-resident arm (value): upload once, run K times
    arm.resident(0, args.warmup, barrier)          # warm-up launches (also builds / loads the specialised kernel)
    if rank == 0:
        sampler.start()
    kernel_ms, wall, res = arm.resident(args.steps, 0, barrier)
    clocks = sampler.stop() if rank == 0 else None
    dev_s_max, wall_max = reduce_max(sum(kernel_ms) * 1e-3, wall)
    total_paths = wg.N * args.steps
    value = total_paths / dev_s_max
    # ---- end-to-end arm
    e2e_dt, r, tm, pin_note = arm.e2e(args.steps, 2, barrier)
    (e2e_dt_max,) = reduce_max(e2e_dt)
    e2e = total_paths / e2e_dt_max
    e2e_same = bool((r.return_code == res.return_code).all() and np.array_equal(r.solution, res.solution))
    # ---- final reduction of the solution-class counts (the only cross-rank exchange of the job)
    stats = reduce_stats(result.statistics(res).asdict())
    peak_gflops = raw.hc_dfma_peak(200000)

    # ---- the other BASELINE.json configs (N = 1) and the strong-scaling C5 sweep (every N)
    per_config, scaling_c5 = {}, None
    if not args.no_per_config:
        if world == 1:
            plan = [("C1 katsura8", "katsura8", dict(replicas=1184)), ("C3 tritangents", "tritangents", {}),
                    ("C4 cyclooctane", "cyclooctane_polyhedral" if os.path.exists(os.path.join(ROOT, "homotopycontinuation.jl_b200", "data", "cyclooctane_cells.json")) else "cyclooctane_td", {}),
                    ("C5 biochem sweep", "biochem_sweep", dict(replicas=512))]
            for label, name, kw in plan:
                try:
                    wc = make_workload(name, kw.get("replicas", 1), api)
                    a2 = Arm(api, raw, wc, opts)
                    a2.resident(0, 1)
                    kms, _, rc = a2.resident(2, 0)
                    dt2, r2, tm2, _ = a2.e2e(1, 1)
                    entry = {"workload": wc.description, "paths_per_step": wc.N, "value": wc.N / (np.mean(kms) * 1e-3), "e2e": wc.N / dt2,
                             "unit": "paths/s", "engine": ENGINES.get(tm2.engine, str(tm2.engine)), "grid": tm2.grid, "block": tm2.block,
                             "class_counts": result.statistics(rc).asdict(), "expected": wc.expected,
                             "roofline_frac": roofline_of(wc, rc, kms, peak_gflops, name, tm2.engine)["frac"]}
                    if tm2.handoff_paths:
                        entry["second_pass"] = {"paths": int(tm2.handoff_paths), "ms": round(tm2.handoff_ms, 1),
                                                "engine": "lane group per path (interpreter): paths beyond 120 endgame steps or in extended precision, tracked again from their starts"}
                    if not args.no_cpu_baseline:
                        budget = {"katsura8": 65536, "tritangents": 16384, "cyclooctane_td": 8192, "cyclooctane_polyhedral": 4096, "biochem_sweep": 524288}[name]
                        v, sample, rcpu, dtc = cpu_arm(wc, budget, os.cpu_count() or 1)
                        entry["cpu"] = {"value": v, "unit": "paths/s", "sample": f"first {sample.N} paths, {dtc:.1f} s", "cores": os.cpu_count() or 1}
                        entry["e2e_over_cpu"] = entry["e2e"] / v
                    per_config[label] = entry
                    del a2, wc
                except Exception as e:   # a config must not take the headline down with it
                    per_config[label] = {"error": repr(e)[:300]}
        # C5 as north_star states it: 10^6 parameter points in total, sharded over the ranks by parameter point
        try:
            import pyoracle   # only to make the generic start solutions once (a total-degree solve of 24 paths)
            from hcb200 import workloads
            p1, starts = workloads.biochem_generic_start(api)
            plo, phi = sharding.shard_range(args.c5_points, rank, world)
            w5 = workloads.biochem_sweep_from_starts(starts, p1, phi - plo, first=plo)
            a5 = Arm(api, raw, w5, opts)
            a5.resident(0, 1, barrier)
            kms5, _, r5 = a5.resident(2, 0, barrier)
            dt5, _, tm5, _ = a5.e2e(2, 1, barrier)
            dev5, e2e5 = reduce_max(sum(kms5) * 1e-3, dt5)
            n5 = args.c5_points * len(starts)
            scaling_c5 = {"workload": f"bio-chemical network 1 parameter sweep, {args.c5_points} parameter points x {len(starts)} start solutions (whole job)",
                          "scaling": "strong", "n_gpus": world, "paths_per_step": n5, "value": 2 * n5 / dev5, "e2e": 2 * n5 / e2e5, "unit": "paths/s",
                          "entry_point": "hc_track_sweep", "class_counts": reduce_stats(result.statistics(r5).asdict())}
            try:   # the same sweep with the results reduced on the device: 20 bytes per point come back (hc_track_sweep_counts)
                from hcb200 import capi as _capi
                _capi.track_sweep_counts(a5.handles["H"], w5.sweep_starts, w5.sweep_q, opts)
                barrier()
                tc0 = time.perf_counter()
                for _ in range(2):
                    cnt5 = _capi.track_sweep_counts(a5.handles["H"], w5.sweep_starts, w5.sweep_q, opts)
                barrier()
                (dtc,) = reduce_max(time.perf_counter() - tc0)
                scaling_c5["e2e_counts_only"] = {"value": 2 * n5 / dtc, "unit": "paths/s", "entry_point": "hc_track_sweep_counts",
                                                 "d2h_bytes_per_step": int(20 * (phi - plo)),
                                                 "nonsingular_real_this_rank": [int(cnt5[:, 0].sum()), int(cnt5[:, 2].sum())]}
            except Exception as e:
                scaling_c5["e2e_counts_only"] = {"error": repr(e)[:200]}
            del a5, w5
        except Exception as e:
            scaling_c5 = {"error": repr(e)[:300]}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": CAPTION, "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wg.description, "paths_per_gpu_per_step": w.N,
                   "parallelism": f"path index range sharded over {world} GPU(s), no collective on the hot path",
                   "l2": f"per-lane path state in local memory, {tm.slab_bytes / 2**20:.0f} MiB over all lanes (> 126 MiB L2 when all lanes run); inputs are KBs",
                   "engine": ENGINES.get(tm.engine, str(tm.engine)),
                   "grid": tm.grid, "block": tm.block, "class_counts": stats, "expected": wg.expected},
        "e2e": {"value": e2e, "unit": "paths/s", "h2d_bytes_per_step": int(tm.h2d_bytes), "d2h_bytes_per_step": int(tm.d2h_bytes),
                "last_call_ms": {"setup_h2d_launch": round(tm.h2d_ms, 2), "kernel": round(tm.kernel_ms, 2), "d2h": round(tm.d2h_ms, 2),
                                 "whole_step_mean": round(1e3 * e2e_dt_max / args.steps, 2)},
                "entry_point": "hc_track_sweep" if w.sweep_starts is not None else (("hc_polyhedral_track_cells" if w.cells is not None else "hc_polyhedral_track_batch") if w.mode == 2 else "hc_track_batch"),
                "host_buffers": pin_note,
                "results_identical_to_resident_arm": e2e_same},
        "gpu_launches": arm.launches,
        "clocks": clocks,
        "roofline": roofline_of(w, res, kernel_ms, peak_gflops, args.workload, tm.engine),
        "wall_s_timed_region": wall_max,
    }
    if not args.no_cpu_baseline and world == 1:
        import pyoracle
        threads = os.cpu_count() or 1
        v, sample, rc, dt = cpu_arm(w, args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": v, "unit": "paths/s", "cores": threads, "kind": "port", "cpu": pyoracle.cpu_model(),
                                "sample": f"first {sample.N} paths of the workload, {threads} std::threads, oracle built -O3 -march=native on this host, {dt:.1f} s"}
    if per_config:
        line["per_config"] = per_config
    if scaling_c5 is not None:
        line["scaling_c5"] = scaling_c5
    print(json.dumps(line))


if __name__ == "__main__":
    main()
