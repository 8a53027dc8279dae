#!/usr/bin/env python
"""bench.py — throughput of the BA hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2p|c4|c5] [--impl ours|reference]

Workloads (BASELINE.json configs; synthetic inputs of rspl_slam_b200/synth.py, seeds of SURVEY §8d)
  c2 (default, configs[1]): batched pose-only FrameOptimization, 4096 frames x 400 stereo points
                            per GPU, Huber + 4 rounds x LM(10)
  c4 (configs[3])         : batched LocalmapOptimization, windows of 10 KF / 3k points / 300 lines,
                            LM 10 + 5, `--windows` per GPU (default 1024)
A "step" is one pass of the hot path (the whole on-device LM schedule) over one batch.

  value  : edges linearised / s, inputs resident in HBM, CUDA events on the solver's stream,
           L2 flushed between steps, max over ranks (units of all ranks / slowest rank's time)
  e2e    : same metric through the C-ABI call with pinned HOST buffers (H2D + solve + D2H timed)
  roofline / cpu_baseline: see DESIGN.md §Measurement
Multi-GPU: one process per GPU (torchrun), independent units sharded by rank, NO data-path
collective -> "scaling": "weak" (per-GPU work fixed).
`--impl reference` times the CPU oracle (the g2o-equivalent restatement of the reference's own
path; g2o itself cannot be built here) on all host threads, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point+line BA edges linearized/sec"
UNIT = "edges/s"
# SURVEY §8(d) contract figures (algorithmic bytes)
BYTES_POSE_ONLY_STEREO = 60  # per stereo edge per evaluation: 24 meas + 24 Xw + 4 flag + 8 chi2
BYTES_POSE_ONLY_MONO = 52
BYTES_POSE_ONLY_STEREO_LINE = 124  # 64 meas + 48 world line + 4 flag + 8 chi2 (same accounting, line extension)
BYTES_POSE_ONLY_MONO_LINE = 92
C2_FRAMES, C2_POINTS, C2_LINES = 4096, 400, 60


def _c2_lines(args):
    """BASELINE configs[1] names 400 stereo points + 60 lines per frame. The reference's FrameOptimization takes no
    lines (g2o_optimization.cc:284-285), so `c2p` (points only) is its parity path and `c2` adds the 60 lines as
    constraints on fixed lines (SURVEY 8a note / 8d)."""
    return 0 if args.workload == "c2p" else args.frame_lines


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self._NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on the host cores (rank 0 only)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return 0
    from oracle import orc
    from rspl_slam_b200 import synth
    threads = orc.max_threads()
    t_all, edges_all, iters_all = [], 0, 0
    if args.workload in ("c2", "c2p"):
        nl = _c2_lines(args)
        sample = min(C2_FRAMES, max(threads * 24, 64))
        desc = f"{sample} of {C2_FRAMES} frames per step (400 stereo pts + {nl} lines, 4x10 LM), {threads} threads, one frame per thread"
        base = [synth.make_frame_problem(synth.config_seed(2, i), n_points=C2_POINTS, n_lines=nl) for i in range(sample)]
        for step in range(args.warmup + args.steps):
            probs = [p.copy() for p in base]
            t0 = time.perf_counter()
            st = orc.frame_opt_batch(probs, n_threads=threads)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                t_all.append(dt)
                edges_all += sum(s["edges_linearized"] for s in st)
                iters_all += sum(sum(s["iters"]) for s in st)
        workload = f"C2 batched pose-only FrameOptimization ({C2_FRAMES} frames x ({C2_POINTS} stereo pts + {nl} lines) per GPU)"
    elif args.workload == "c5":
        # one problem: the oracle is single-threaded per problem and factorises densely, so the bounded sample is a
        # scaled-down problem of the same generator
        desc = "scaled-down C5 (60 KF / 30k points / 3k lines, same generator) per step, 1 thread (one problem)"
        threads = 1
        base = [synth.make_global_problem(synth.config_seed(5, 0), n_kf=60, n_points=30000, n_lines=3000, loops=1)]
        for step in range(args.warmup + args.steps):
            probs = [p.copy() for p in base]
            t0 = time.perf_counter()
            st = orc.local_ba_batch(probs, n_threads=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                t_all.append(dt)
                edges_all += sum(s["edges_linearized"] for s in st)
                iters_all += sum(sum(s["iters"]) for s in st)
        workload = f"C5 global BA ({args.kf} KF / {args.points} points / {args.lines} lines)"
    else:
        sample = max(threads, 8)
        desc = f"{sample} of {args.windows} windows per step (10 KF/3k pts/300 lines, LM 10+5), {threads} threads"
        base = [synth.make_local_problem(synth.config_seed(4, i)) for i in range(sample)]
        for step in range(args.warmup + args.steps):
            probs = [p.copy() for p in base]
            t0 = time.perf_counter()
            st = orc.local_ba_batch(probs, n_threads=threads)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                t_all.append(dt)
                edges_all += sum(s["edges_linearized"] for s in st)
                iters_all += sum(sum(s["iters"]) for s in st)
        workload = f"C4 batched LocalmapOptimization ({args.windows} windows per GPU)"
    total = float(sum(t_all))
    value = edges_all / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(t_all), 1), "higher_is_better": True,
        "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "reference_impl": "CPU oracle (g2o-equivalent restatement; g2o/Eigen not installable here)"},
        "lm_iters_per_sec": iters_all / total,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _pin_batch(batch, capi):
    """Copies every array of a batch into page-locked host memory (e2e copies come from pinned memory)."""
    kw = {}
    for k, v in batch.__dict__.items():
        if isinstance(v, np.ndarray):
            p = capi.pinned_empty(v.shape, v.dtype)
            p[...] = v
            kw[k] = p
        else:
            kw[k] = v
    return type(batch)(**kw)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rspl_slam_b200 import capi, synth

    rank, local_rank, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    ctx = capi.Context(device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    opt = capi.make_options()

    # ---- inputs: this rank's shard of independent units (weak scaling: fixed work per GPU)
    if args.workload in ("c2", "c2p"):
        n_units = args.frames
        nl = _c2_lines(args)
        batch = synth.make_frame_batch(2, n_units, first_instance=rank * n_units, n_points=C2_POINTS, n_lines=nl)
        upload, solve = ctx.frame_batch_upload, ctx.frame_batch_solve
        out = ctx.alloc_frame_result(batch, pinned=True)
        download = lambda: ctx.frame_batch_download(out)
        oneshot = lambda b: ctx.frame_batch(b, opt, out)
        workload = (f"C2 batched pose-only FrameOptimization ({n_units} frames x ({C2_POINTS} stereo pts + {nl} lines) per GPU, "
                    "Huber + 4 rounds x LM10" + ("; lines = constraints on fixed 3-D lines, an extension: the reference's "
                    "FrameOptimization takes none" if nl else "; points only = the reference's FrameOptimization") + ")")
        kernel = "ba::frame_opt_kernel"
    else:
        n_units = args.windows
        batch, _ = synth.make_local_batch(4, n_units, first_instance=rank * n_units)
        upload, solve = ctx.local_batch_upload, ctx.local_batch_solve
        out = ctx.alloc_local_result(batch, pinned=True)
        download = lambda: ctx.local_batch_download(out)
        oneshot = lambda b: ctx.local_batch(b, opt, out)
        workload = f"C4 batched LocalmapOptimization ({n_units} windows of 10 KF/3k pts/300 lines per GPU, LM 10+5)"
        kernel = "ba::local_ba_kernel"
    is_c2 = args.workload in ("c2", "c2p")
    if is_c2 and len(batch.cameras) == 1:
        # one camera and all ->inlier flags true are the ABI defaults: pass NULL instead of copying 8 MB of zeros / ones
        batch.mono_cam = batch.stereo_cam = None
        if batch.mono_inlier.all() and batch.stereo_inlier.all():
            batch.mono_inlier = batch.stereo_inlier = None
        if batch.mline_begin is not None:
            batch.mline_cam = batch.sline_cam = None
            if batch.mline_inlier.all() and batch.sline_inlier.all():
                batch.mline_inlier = batch.sline_inlier = None
    pinned = _pin_batch(batch, capi)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-input timing
    upload(pinned)
    launches0 = ctx.launch_count
    for _ in range(max(args.warmup, 3)):
        solve(opt)
    ctx.sync()
    download()
    stats = out.stats.copy()
    warm_launches = ctx.launch_count - launches0
    launches_per_step = warm_launches // max(args.warmup, 3)

    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    with ClockSampler(local_rank) as clk:
        t_wall0 = time.perf_counter()
        with torch.cuda.stream(stream):
            for i in range(args.steps):
                flush.zero_()  # evict L2 between timed steps (not timed)
                starts[i].record(stream)
                solve(opt)
                ends[i].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    dev_s = float(sum(step_ms)) * 1e-3
    gpu_launches = launches_per_step * args.steps

    edges_lin = int(stats["edges_linearized"].sum())
    edges_eval = int(stats["edges_evaluated"].sum())
    lm_iters = int(stats["iters"].sum())
    lm_trials = int(stats["trials"].sum())

    # ---- end to end through the C-ABI with pinned host buffers
    for _ in range(2):
        oneshot(pinned)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        oneshot(pinned)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    t = torch.tensor([dev_s, e2e_s, t_wall], dtype=torch.float64, device=dev)
    tot = torch.tensor([edges_lin, lm_iters, edges_eval, lm_trials], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_s_max, e2e_s_max, wall_max = [float(x) for x in t.tolist()]
    edges_lin_all, lm_iters_all, edges_eval_all, lm_trials_all = [float(x) for x in tot.tolist()]

    if rank == 0:
        value = edges_lin_all * args.steps / dev_s_max
        e2e_value = edges_lin_all * e2e_steps / e2e_s_max
        peak, peak_src = _peaks()
        # roofline of the dominant kernel class: algorithmic bytes / CUDA-event time of
        # that class, measured in a separate profiled pass (event pairs around every launch)
        prof_steps = max(1, min(args.steps, 5))
        ctx.set_profiling(True)
        for _ in range(prof_steps):
            solve(opt)
        prof = ctx.get_profile()
        ctx.set_profiling(False)
        per_kernel = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps}
                      for k, v in prof.items() if v[1]}
        if is_c2:
            cnt = {BYTES_POSE_ONLY_STEREO: int(batch.stereo_begin[-1]), BYTES_POSE_ONLY_MONO: int(batch.mono_begin[-1])}
            if batch.mline_begin is not None:
                cnt[BYTES_POSE_ONLY_STEREO_LINE] = int(batch.sline_begin[-1])
                cnt[BYTES_POSE_ONLY_MONO_LINE] = int(batch.mline_begin[-1])
            # evaluations are spread over the edge classes in proportion to their counts
            alg_bytes = edges_eval * sum(bts * n for bts, n in cnt.items()) / max(sum(cnt.values()), 1)
            dom_ms = prof["frame_opt"][0] / prof_steps
            dom_launches = prof["frame_opt"][1] / prof_steps
            dom_name = kernel
        else:
            from rspl_slam_b200.roofline import local_class_bytes
            cls_bytes = local_class_bytes(batch, stats)  # bytes per step of the linearise / Schur / back-sub classes
            groups = {"linearize (kb_linearize + kb_pose_blocks)": (("linearize", "pose_blocks"), cls_bytes["linearize"]),
                      "schur (kt_schur_tile | kb_schur_prep + kb_schur_reduce, + kb_solve)": (("schur_tile", "schur_prep", "schur_reduce", "reduced_solve"), cls_bytes["schur"]),
                      "backsub (kb_backsub)": (("backsub_update_eval",), cls_bytes["backsub"]),
                      "persistent (local_solve_kernel)": (("local_solve_persistent",), sum(cls_bytes.values()))}
            best = None
            for name, (classes, nbytes) in groups.items():
                ms = sum(prof[c][0] for c in classes) / prof_steps
                nl = sum(prof[c][1] for c in classes) / prof_steps
                if nl == 0:
                    continue
                per_kernel[name] = {"ms_per_step": ms, "algorithmic_bytes_per_step": nbytes,
                                    "achieved_GBs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
                if best is None or ms > best[1]:
                    best = (name, ms, nl, nbytes)
            dom_name, dom_ms, dom_launches, alg_bytes = best
        launch_s = dom_ms * 1e-3 / max(dom_launches, 1)
        alg_per_launch = alg_bytes / max(dom_launches, 1)
        achieved = alg_per_launch / launch_s / 1e9
        cpu = cpu_baseline(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2": "flushed (256 MiB write) between timed steps; inputs resident in HBM",
                       "parallelism": f"dp{world} (independent units, no collective)", "timing": "CUDA events on the solver stream, max over ranks"},
            "lm_iters_per_sec": lm_iters_all * args.steps / dev_s_max,
            "lm_trials_per_sec": lm_trials_all * args.steps / dev_s_max,
            "edges_evaluated_per_sec": edges_eval_all * args.steps / dev_s_max,
            "units_per_sec": n_units * world * args.steps / dev_s_max,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pinned.h2d_bytes()),
                    "d2h_bytes_per_step": int(out.d2h_bytes()), "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                    "api": "rspl_ba_%s_batch (pinned host buffers in, pinned host buffers out)" % ("frame" if is_c2 else "local")},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": 1e3 * launch_s,
                         "launches_per_step": dom_launches, "per_kernel": per_kernel,
                         "note": "achieved = SURVEY 8(d) contract bytes (materialised-W formulation) / CUDA-event time of the kernel class; "
                                 "the kernels recompute instead of materialising, so real DRAM traffic (`traffic`) is lower, see DESIGN.md"},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
            "wall_s_timed_region": wall_max,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


# ------------------------------------------------------------------------------------------------
# C5: ONE global problem, landmarks partitioned over the ranks (strong scaling; NCCL all-reduces of the
# pose blocks, the Schur complement pieces and a few scalars per LM trial)
# ------------------------------------------------------------------------------------------------
def run_global(args):
    import torch
    import torch.distributed as dist
    from rspl_slam_b200 import capi, synth
    from rspl_slam_b200.problem import LocalBatch, shard_landmarks
    from rspl_slam_b200.roofline import local_class_bytes

    rank, local_rank, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    opt = capi.make_options()
    ident = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if world > 1:
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, src=0)
    ctx.comm_init(world, rank, bytes(ident.cpu().numpy().tobytes()) if world > 1 else None)

    full = synth.make_global_problem(synth.config_seed(5, 0), n_kf=args.kf, n_points=args.points, n_lines=args.lines, loops=3)
    shard = shard_landmarks(full, rank, world)
    batch = LocalBatch.from_problems([shard.problem])
    pinned = _pin_batch(batch, capi)
    out = ctx.alloc_local_result(batch, pinned=True)
    workload = (f"C5 global BA: {args.kf} KF / {len(full.point_id)} points / {len(full.line_id)} lines / {full.n_edges} constraints, "
                f"LM 10+5, landmarks partitioned over {world} GPU(s)")
    total_edges = full.n_edges
    del full
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 1)
    ctx.global_upload(pinned)
    launches0, coll0 = ctx.launch_count, ctx.collective_count()
    for _ in range(warm):
        ctx.global_solve(opt)
    ctx.sync()
    ctx.global_download(out)
    stats = out.stats.copy()
    launches_per_step = (ctx.launch_count - launches0) // warm
    coll_per_step = (ctx.collective_count() - coll0) // warm
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    with ClockSampler(local_rank) as clk:
        t_wall0 = time.perf_counter()
        with torch.cuda.stream(stream):
            for i in range(args.steps):
                flush.zero_()
                starts[i].record(stream)
                ctx.global_solve(opt)
                ends[i].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    dev_s = float(sum(s.elapsed_time(e) for s, e in zip(starts, ends))) * 1e-3
    # end to end: host buffers in, host buffers out
    ctx.global_ba(pinned, opt)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.global_upload(pinned)
        ctx.global_solve(opt)
        ctx.global_download(out)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([dev_s, e2e_s, t_wall], dtype=torch.float64, device=dev)
    byt = torch.tensor([float(pinned.h2d_bytes()), float(out.d2h_bytes())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(byt, op=dist.ReduceOp.SUM)
    dev_s_max, e2e_s_max, wall_max = [float(x) for x in t.tolist()]
    # the LM statistics are global (every rank holds the same numbers)
    edges_lin, edges_eval = int(stats["edges_linearized"][0]), int(stats["edges_evaluated"][0])
    lm_iters, lm_trials = int(stats["iters"][0].sum()), int(stats["trials"][0].sum())
    if rank == 0:
        peak, peak_src = _peaks()
        ctx.set_profiling(True)
    # (the profiled pass is collective too: every rank runs it, rank 0 records events)
    ctx.global_solve(opt)
    if rank == 0:
        prof = ctx.get_profile()
        ctx.set_profiling(False)
        per_kernel = {k: {"ms_per_step": v[0], "launches_per_step": v[1]} for k, v in prof.items() if v[1]}
        # stats count the linearised edges of ALL ranks; this rank's kernels touched its shard's share of them
        stats_local = stats.copy()
        stats_local["edges_linearized"] = (stats["edges_linearized"] * (batch.n_edges / max(total_edges, 1))).astype(stats["edges_linearized"].dtype)
        cls_bytes = local_class_bytes(batch, stats_local)  # this rank's share of the contract bytes
        groups = {"linearize (kb_linearize + kb_pose_blocks)": (("linearize", "pose_blocks"), cls_bytes["linearize"]),
                  "schur (kb_schur_prep + kb_schur_reduce + dense solve)": (("schur_prep", "schur_reduce", "reduced_solve"), cls_bytes["schur"]),
                  "backsub (kb_backsub)": (("backsub_update_eval",), cls_bytes["backsub"])}
        best = None
        for name, (classes, nbytes) in groups.items():
            ms = sum(prof[c][0] for c in classes)
            nl = sum(prof[c][1] for c in classes)
            if nl == 0:
                continue
            per_kernel[name] = {"ms_per_step": ms, "algorithmic_bytes_per_step": nbytes,
                                "achieved_GBs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
            if best is None or ms > best[1]:
                best = (name, ms, nl, nbytes)
        dom_name, dom_ms, dom_launches, alg_bytes = best
        launch_s = dom_ms * 1e-3 / max(dom_launches, 1)
        line = {
            "metric": METRIC, "value": edges_lin * args.steps / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2": "flushed (256 MiB write) between timed steps; inputs resident in HBM",
                       "parallelism": f"landmark-sharded x{world}, poses replicated, {coll_per_step} NCCL all-reduces per solve, "
                                      "replicated dense Cholesky of the reduced camera system",
                       "timing": "CUDA events on the solver stream, max over ranks"},
            "lm_iters_per_sec": lm_iters * args.steps / dev_s_max, "lm_trials_per_sec": lm_trials * args.steps / dev_s_max,
            "edges_evaluated_per_sec": edges_eval * args.steps / dev_s_max, "units_per_sec": args.steps / dev_s_max,
            "e2e": {"value": edges_lin * e2e_steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(byt[0].item()),
                    "d2h_bytes_per_step": int(byt[1].item()), "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                    "api": "rspl_ba_global_upload / _solve / _download (pinned host buffers, all ranks)"},
            "gpu_launches": int(launches_per_step * args.steps), "collectives_per_step": int(coll_per_step),
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": alg_bytes / max(dom_launches, 1) / launch_s / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg_bytes / max(dom_launches, 1) / launch_s / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / max(dom_launches, 1),
                         "launch_ms": 1e3 * launch_s, "launches_per_step": dom_launches, "per_kernel": per_kernel,
                         "note": "rank 0's kernels on its landmark shard; contract bytes of SURVEY 8(d)"},
            "cpu_baseline": cpu_baseline(args), "clocks": clk.summary(), "wall_s_timed_region": wall_max,
        }
        print(json.dumps(line), flush=True)
    barrier()
    ctx.comm_destroy()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


def _traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def cpu_baseline(args):
    """The oracle timed on this box's host cores on a bounded sample (1 thread: default g2o and the
    reference's call pattern are single-threaded)."""
    from oracle import orc
    from rspl_slam_b200 import synth
    if args.workload in ("c2", "c2p"):
        n = 1024 if args.workload == "c2p" else 256  # (g2o differentiates line edges numerically: ~6x the work per frame)
        probs = [synth.make_frame_problem(synth.config_seed(2, i), n_points=C2_POINTS, n_lines=_c2_lines(args)) for i in range(n)]
        t0 = time.perf_counter()
        st = orc.frame_opt_batch(probs, n_threads=1)
        dt = time.perf_counter() - t0
        sample = f"first {n} of {C2_FRAMES} C2 frames ({C2_POINTS} stereo pts + {_c2_lines(args)} lines), 1 thread"
    elif args.workload == "c5":
        # the oracle factorises the reduced system densely in one thread: a scaled-down problem of the same generator
        probs = [synth.make_global_problem(synth.config_seed(5, 0), n_kf=60, n_points=30000, n_lines=3000, loops=1)]
        t0 = time.perf_counter()
        st = orc.local_ba_batch(probs, n_threads=1)
        dt = time.perf_counter() - t0
        sample = "scaled-down C5 (60 KF / 30k points / 3k lines, same generator), 1 thread"
    else:
        n = 8
        probs = [synth.make_local_problem(synth.config_seed(4, i)) for i in range(n)]
        t0 = time.perf_counter()
        st = orc.local_ba_batch(probs, n_threads=1)
        dt = time.perf_counter() - t0
        sample = f"first {n} C4 windows, 1 thread"
    edges = sum(s["edges_linearized"] for s in st)
    return {"value": edges / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
            "seconds": dt, "lm_iters_per_sec": sum(sum(s["iters"]) for s in st) / dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c2p", "c4", "c5"],
                    help="c2: pose-only frames, 400 stereo points + 60 lines (BASELINE configs[1]); c2p: points only (the reference's path)")
    ap.add_argument("--frame-lines", type=int, default=C2_LINES, help="lines per frame (c2)")
    ap.add_argument("--kf", type=int, default=2000, help="keyframes of the global problem (c5)")
    ap.add_argument("--points", type=int, default=1_000_000, help="points of the global problem (c5)")
    ap.add_argument("--lines", type=int, default=100_000, help="lines of the global problem (c5)")
    ap.add_argument("--frames", type=int, default=C2_FRAMES, help="frames per GPU (c2)")
    ap.add_argument("--windows", type=int, default=1024, help="windows per GPU (c4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_global(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
