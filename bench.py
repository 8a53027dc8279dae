#!/usr/bin/env python
"""bench.py — throughput of the BA hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c2|c2p|c4|c1|c3|frame1|c5] [--impl ours|reference]

ONE JSON line. With `--workload all` (default) the headline keys are C2 (BASELINE configs[1], the configuration
the metric is quoted on) and `workloads` holds one sub-object per further configuration, each with its own
`value`, `ms_per_step`, `e2e`, `roofline`, `cpu_baseline`:
  c2     configs[1]: batched pose-only FrameOptimization, 4096 frames x (400 stereo points + 60 lines) per GPU,
         Huber + 4 rounds x LM(10). The 60 lines are an extension (constraints on fixed lines, SURVEY 8a note).
  c2p    the same, points only = the reference's own FrameOptimization (g2o_optimization.cc:256-397)
  c4     configs[3]: 1024 LocalmapOptimization windows (10 KF / 3k points / 300 lines) per GPU, LM 10 + 5
  c1     configs[0]: ONE such window (the reference's call pattern, map.cc:709-710): latency
  c3     configs[2]: one 20-KF / 10k-point / 1k-line window: latency
  frame1 one FrameOptimization call on one frame (map_builder.cc:583-584): latency
  c5     configs[4]: one global BA, 2000 KF / 1M points / 100k lines, landmarks partitioned over the ranks, NCCL
         all-reduces from the library's own communicator (strong scaling); at N >= 2 a `parity_check` on a
         test-scale problem (poses bit-identical on all ranks, and equal to the un-sharded solve within tolerance)
A "step" is one pass of the hot path (the whole on-device LM schedule) over one batch.
  value  : edges linearised / s, inputs resident in HBM, CUDA events on the solver's stream, L2 flushed between
           steps, max over ranks (units of all ranks / slowest rank's time)
  e2e    : same metric through the C-ABI call with pinned HOST buffers (H2D + solve + D2H timed)
  roofline / cpu_baseline: see DESIGN.md §Measurement
Multi-GPU: one process per GPU (torchrun). C2 / C4: independent units sharded by rank, no data-path collective
("weak"); C5: one problem, "strong". c1 / c3 / frame1 do not shard (replicas only): N = 1 runs only.
`--impl reference` times the CPU oracle (the g2o-equivalent restatement of the reference's own path; g2o itself
cannot be built here) on all host threads, rank 0 only, for the same workloads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
ALL_N1 = ["c2", "c2p", "c4", "c1", "c3", "frame1", "c5", "tri", "ends"]
ALL_MULTI = ["c2", "c2p", "c4", "c5"]


def workload_name(key, args):
    """config.workload — the same string in both arms."""
    if key == "c2":
        return (f"C2 batched pose-only FrameOptimization: {args.frames} frames x ({C2_POINTS} stereo points + "
                f"{args.frame_lines} lines) per GPU, Huber + 4 rounds x LM10")
    if key == "c2p":
        return (f"C2p batched pose-only FrameOptimization, points only (the reference's own path): {args.frames} frames x "
                f"{C2_POINTS} stereo points per GPU, Huber + 4 rounds x LM10")
    if key == "c4":
        return f"C4 batched LocalmapOptimization: {args.windows} windows of 10 KF / 3k points / 300 lines per GPU, LM 10+5"
    if key == "c1":
        return "C1 one LocalmapOptimization window: 10 KF / 3k points / 300 lines, LM 10+5 (latency)"
    if key == "c3":
        return "C3 one LocalmapOptimization window: 20 KF / 10k points / 1k lines, LM 10+5 (latency)"
    if key == "frame1":
        return f"one FrameOptimization call: 1 frame x {C2_POINTS} stereo points, Huber + 4 rounds x LM10 (latency)"
    if key == "c5":
        return (f"C5 global BA: {args.kf} KF / {args.points} points / {args.lines} lines on a 3-loop trajectory, LM 10+5, "
                "landmarks partitioned over the GPUs" + (f", loop closures every {args.closures} blocks" if getattr(args, "closures", 0) else ""))
    if key == "tri":
        return (f"TRI batched Map::TriangulateMappoint (SURVEY 8f-4): {TRI_POINTS} new map points x 0-8 observations from "
                f"{TRI_FRAMES} keyframes")
    if key == "ends":
        return (f"ENDS batched Map::UppdateMapline (SURVEY 8f-2): endpoint refresh of {ENDS_LINES} optimised lines "
                f"(1024 windows x 300) x 0-{ENDS_MAX_PTS} map points")
    raise ValueError(key)


ENDS_LINES, ENDS_MAX_PTS, ENDS_SAMPLE = 307_200, 24, 307_200
ENDS_METRIC, ENDS_UNIT = "map lines refreshed/sec", "lines/s"


def _ends_batch(n_lines):
    from rspl_slam_b200 import synth
    return synth.make_mapline_batch(20261018 + 7000, n_lines=n_lines, max_pts=ENDS_MAX_PTS)


def _pin_arrays(b, capi):
    """the numpy arrays of a dict copied into page-locked host memory"""
    out = {}
    for k, v in b.items():
        if isinstance(v, np.ndarray) and v.size:
            p = capi.pinned_empty(v.shape, v.dtype)
            p[...] = v
            out[k] = p
        else:
            out[k] = v
    return out


def bench_ends(env, steps, warmup, with_cpu):
    """SURVEY 8(f) rank 2. value: lines per second of the kernel alone (event pair around the launch, inputs
    resident); e2e: the same through rspl_ba_update_maplines with host arrays (copies inside)."""
    from oracle import orc
    ctx, args = env.ctx, env.args
    b = _ends_batch(ENDS_LINES)
    n, n_ref, n_pts = len(b["pt_begin"]) - 1, int(b["pt_begin"][-1]), b["point_xyz"].shape[1]
    from rspl_slam_b200 import capi
    b = _pin_arrays(b, capi)  # the e2e copies come from / go to page-locked host memory
    out = (capi.pinned_empty((6, n), np.float64), capi.pinned_empty((n,), np.uint8))
    out[0][:] = 0.0
    call = lambda: ctx.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"], out=out)
    for _ in range(warmup):
        call()
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    with ClockSampler(env.local_rank) as clk:
        for _ in range(steps):
            env.flush.zero_()
            env.torch.cuda.synchronize(env.dev)
            ends, ok, cnt = call()
    wall = time.perf_counter() - t0
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    ms_kernel = prof["frame_opt"][0] / max(prof["frame_opt"][1], 1)
    t1 = time.perf_counter()
    for _ in range(steps):
        call()
    e2e_s = (time.perf_counter() - t1) / steps
    # per line: Line3D (48) + offset (4) + endpoints (48) + flag (1); per reference: index (4) + position (24, each point
    # is referenced once on average, so the point array is read once)
    alg_bytes = 101.0 * n + 28.0 * n_ref
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    res = {
        "metric": ENDS_METRIC, "value": n / (ms_kernel * 1e-3), "unit": ENDS_UNIT, "n_gpus": env.world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("ends", args), "l2": "flushed (256 MiB write) between timed steps",
                   "timing": "CUDA event pair around the kernel (profiling class of the C-ABI call)"},
        "lines_ok": int(cnt), "point_references": n_ref,
        "e2e": {"value": n / e2e_s, "unit": ENDS_UNIT, "h2d_bytes_per_step": int(100 * n + 4 + 4 * n_ref + 24 * n_pts),
                "d2h_bytes_per_step": int(49 * n + 4), "ms_per_step": 1e3 * e2e_s,
                "api": "rspl_ba_update_maplines (pinned host arrays in, pinned host arrays out)"},
        "gpu_launches": int(prof["frame_opt"][1]),
        "roofline": {"bound": "hbm", "kernel": "ba::line_endpoints_kernel", "achieved": achieved, "peak": env.peak,
                     "unit": "GB/s", "frac": achieved / env.peak, "traffic": None, "peak_source": env.peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": ms_kernel},
        "clocks": clk.summary(), "wall_s_timed_region": wall,
    }
    if with_cpu:
        t2 = time.perf_counter()
        ref_ends, ref_ok, _ = orc.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"])
        dt = time.perf_counter() - t2
        res["cpu_baseline"] = {"value": n / dt, "unit": ENDS_UNIT, "cores": 1, "kind": "port",
                               "sample": f"all {n} lines, 1 thread", "seconds": dt}
        res["parity_check"] = bool(np.array_equal(ref_ok, ok) and np.array_equal(ref_ends, ends))
    return res


TRI_POINTS, TRI_FRAMES = 1_000_000, 16
TRI_METRIC, TRI_UNIT = "map points triangulated/sec", "points/s"


def _tri_batch(n_points):
    """the generator is a python loop: a 20 k-point sample is tiled to the benchmark size (same content per tile)"""
    from rspl_slam_b200 import synth
    b = synth.make_triangulation_batch(20261018 + 6000, n_points=min(n_points, 20000), n_frames=TRI_FRAMES, max_obs=8)
    reps = (n_points + len(b["obs_begin"]) - 2) // (len(b["obs_begin"]) - 1)
    if reps > 1:
        nobs = np.diff(b["obs_begin"])
        nobs_t = np.tile(nobs, reps)[:n_points]
        n_full = int(nobs_t.sum())
        b = dict(b, obs_begin=np.concatenate([[0], np.cumsum(nobs_t)]).astype(np.int32),
                 obs_frame=np.tile(b["obs_frame"], reps)[:n_full].copy(),
                 obs_uv=np.ascontiguousarray(np.tile(b["obs_uv"], (1, reps))[:, :n_full]))
    return b


def bench_tri(env, steps, warmup, with_cpu):
    """SURVEY 8(f) rank 4. value: points per second of the kernel alone (event pair around the launch, inputs
    resident); e2e: the same through rspl_ba_triangulate_points with host arrays (copies inside)."""
    from oracle import orc
    ctx, args = env.ctx, env.args
    b = _tri_batch(TRI_POINTS)
    n, n_obs = len(b["obs_begin"]) - 1, int(b["obs_begin"][-1])
    from rspl_slam_b200 import capi
    b = _pin_arrays(b, capi)  # the e2e copies come from / go to page-locked host memory
    out = (capi.pinned_empty((3, n), np.float64), capi.pinned_empty((n,), np.uint8))
    out[0][:] = 0.0
    call = lambda: ctx.triangulate_points(b["obs_begin"], b["obs_frame"], b["obs_uv"], b["frame_twc"], b["cam5"], out=out)
    for _ in range(warmup):
        call()
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    with ClockSampler(env.local_rank) as clk:
        for _ in range(steps):
            env.flush.zero_()
            env.torch.cuda.synchronize(env.dev)
            xyz, ok, cnt = call()
    wall = time.perf_counter() - t0
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    ms_kernel = prof["frame_opt"][0] / max(prof["frame_opt"][1], 1)
    t1 = time.perf_counter()
    for _ in range(steps):
        call()
    e2e_s = (time.perf_counter() - t1) / steps
    alg_bytes = 20.0 * n_obs + 29.0 * n  # frame index + pixel (20 B) per observation; offsets + position + flag per point
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    res = {
        "metric": TRI_METRIC, "value": n / (ms_kernel * 1e-3), "unit": TRI_UNIT, "n_gpus": env.world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("tri", args), "l2": "flushed (256 MiB write) between timed steps",
                   "timing": "CUDA event pair around the kernel (profiling class of the C-ABI call)"},
        "points_ok": int(cnt), "observations": n_obs,
        "e2e": {"value": n / e2e_s, "unit": TRI_UNIT, "h2d_bytes_per_step": int(20 * n_obs + 28 * n + 4 + 56 * TRI_FRAMES),
                "d2h_bytes_per_step": int(25 * n), "ms_per_step": 1e3 * e2e_s,
                "api": "rspl_ba_triangulate_points (pinned host arrays in, pinned host arrays out)"},
        "gpu_launches": int(prof["frame_opt"][1]),
        "roofline": {"bound": "hbm", "kernel": "ba::triangulate_points_kernel", "achieved": achieved, "peak": env.peak,
                     "unit": "GB/s", "frac": achieved / env.peak, "traffic": None, "peak_source": env.peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": ms_kernel},
        "clocks": clk.summary(), "wall_s_timed_region": wall,
    }
    if with_cpu:
        sample = _tri_batch(200_000)
        t2 = time.perf_counter()
        orc.triangulate_points(sample["obs_begin"], sample["obs_frame"], sample["obs_uv"], sample["frame_twc"], sample["cam5"])
        dt = time.perf_counter() - t2
        res["cpu_baseline"] = {"value": 200_000 / dt, "unit": TRI_UNIT, "cores": 1, "kind": "port",
                               "sample": "200000 of 1000000 points, 1 thread", "seconds": dt}
    return res


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _fp64_peak():
    """FP64 FMA peak measured by profiles/scripts/fp64_peak.py on this pool's B200 (committed), else the vendor figure."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["fp64_tflops"]), "measured (profiles/fp64_peak.json)"
        except Exception:
            pass
    return 37.0, "vendor figure (not measured)"


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


def _local_kw(key):
    return dict(n_kf=20, n_points=10000, n_lines=1000) if key == "c3" else {}


# ------------------------------------------------------------------------------------------------
# CPU oracle legs (reference arm and cpu_baseline): bounded samples of the same workloads
# ------------------------------------------------------------------------------------------------
def _oracle_sample(key, args, threads, budget):
    """Returns (problems, runner, description): a bounded sample of workload `key` for `threads` oracle threads,
    sized for roughly `budget` seconds of CPU work."""
    from oracle import orc
    from rspl_slam_b200 import synth
    if key in ("c2", "c2p", "frame1"):
        nl = args.frame_lines if key == "c2" else 0
        per_frame = 0.12 if nl else 0.02  # s per frame per thread (g2o differentiates line edges numerically)
        n = 1 if key == "frame1" else int(min(args.frames, max(threads, budget * threads / per_frame)))
        probs = [synth.make_frame_problem(synth.config_seed(2, i), n_points=C2_POINTS, n_lines=nl) for i in range(n)]
        desc = (f"{n} of {args.frames} frames per step ({C2_POINTS} stereo pts + {nl} lines, 4x10 LM), {threads} thread(s), "
                "one frame per thread")
        return probs, (lambda ps: orc.frame_opt_batch(ps, n_threads=threads)), desc
    if key in ("c4", "c1", "c3"):
        cfg = {"c4": 4, "c1": 1, "c3": 3}[key]
        n = 1 if key != "c4" else int(min(args.windows, max(threads, budget * threads / 0.35)))
        probs = [synth.make_local_problem(synth.config_seed(cfg, i), **_local_kw(key)) for i in range(n)]
        desc = (f"{n} of {args.windows} windows per step, {threads} thread(s), one window per thread" if key == "c4"
                else "the whole window, 1 thread (one problem)")
        return probs, (lambda ps: orc.local_ba_batch(ps, n_threads=threads if key == "c4" else 1)), desc
    if key == "c5":
        # one problem: the oracle is single-threaded per problem and factorises densely, so the bounded sample is a
        # scaled-down problem of the same generator
        probs = [synth.make_global_problem(synth.config_seed(5, 0), n_kf=60, n_points=30000, n_lines=3000, loops=1)]
        return probs, (lambda ps: orc.local_ba_batch(ps, n_threads=1)), \
            "scaled-down C5 (60 KF / 30k points / 3k lines, same generator), 1 thread (one problem)"
    raise ValueError(key)


def cpu_baseline(key, args, budget=6.0):
    """The oracle timed on this box's host cores on a bounded sample (1 thread: default g2o and the reference's
    call pattern are single-threaded)."""
    probs, runner, desc = _oracle_sample(key, args, 1, budget)
    t0 = time.perf_counter()
    st = runner(probs)
    dt = time.perf_counter() - t0
    edges = sum(s["edges_linearized"] for s in st)
    return {"value": edges / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "seconds": dt,
            "lm_iters_per_sec": sum(sum(s["iters"]) for s in st) / dt}


def reference_tri(args, steps, warmup):
    from oracle import orc
    sample = _tri_batch(200_000)
    ts = []
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        orc.triangulate_points(sample["obs_begin"], sample["obs_frame"], sample["obs_uv"], sample["frame_twc"], sample["cam5"])
        if step >= warmup:
            ts.append(time.perf_counter() - t0)
    value = 200_000 * len(ts) / float(sum(ts))
    return {"impl": "reference", "metric": TRI_METRIC, "value": value, "unit": TRI_UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * float(sum(ts)) / len(ts), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name("tri", args),
                       "reference_impl": "CPU oracle (restatement of map.cc:292-339 + Eigen's ColPivHouseholderQR)"},
            "cpu_baseline": {"value": value, "unit": TRI_UNIT, "cores": 1, "kind": "port", "sample": "200000 of 1000000 points, 1 thread"},
            "e2e": {"value": value, "unit": TRI_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def reference_ends(args, steps, warmup):
    from oracle import orc
    b = _ends_batch(ENDS_SAMPLE)
    ts = []
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        orc.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"])
        if step >= warmup:
            ts.append(time.perf_counter() - t0)
    value = ENDS_SAMPLE * len(ts) / float(sum(ts))
    return {"impl": "reference", "metric": ENDS_METRIC, "value": value, "unit": ENDS_UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * float(sum(ts)) / len(ts), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name("ends", args),
                       "reference_impl": "CPU oracle (restatement of map.cc:121-177 + g2o's Line3D::toCartesian with Eigen's LDLT)"},
            "cpu_baseline": {"value": value, "unit": ENDS_UNIT, "cores": 1, "kind": "port", "sample": f"all {ENDS_SAMPLE} lines, 1 thread"},
            "e2e": {"value": value, "unit": ENDS_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def reference_one(key, args, steps, warmup, budget):
    from oracle import orc
    if key == "tri":
        return reference_tri(args, steps, warmup)
    if key == "ends":
        return reference_ends(args, steps, warmup)
    threads = orc.max_threads()
    use = threads if key in ("c2", "c2p", "c4") else 1
    base, runner, desc = _oracle_sample(key, args, use, budget)
    t_all, edges_all, iters_all = [], 0, 0
    for step in range(warmup + steps):
        probs = [p.copy() for p in base]
        t0 = time.perf_counter()
        st = runner(probs)
        dt = time.perf_counter() - t0
        if step >= warmup:
            t_all.append(dt)
            edges_all += sum(s["edges_linearized"] for s in st)
            iters_all += sum(sum(s["iters"]) for s in st)
    total = float(sum(t_all))
    value = edges_all / total
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / max(len(t_all), 1), "higher_is_better": True,
        "scaling": "strong" if key == "c5" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(key, args),
                   "reference_impl": "CPU oracle (g2o-equivalent restatement; g2o/Eigen not installable here)"},
        "lm_iters_per_sec": iters_all / total,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": use, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return 0
    keys = (ALL_N1 if args.gpus <= 1 else ALL_MULTI) if args.workload == "all" else [args.workload]
    line = reference_one(keys[0], args, args.steps, args.warmup, budget=2.0)
    if len(keys) > 1:
        line["workloads"] = {k: reference_one(k, args, max(1, min(args.steps, 2)), 1, budget=1.5) for k in keys[1:]}
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


def _traffic(key):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", f"traffic_{key}.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def _bind_near_gpu(index):
    """Binds this process to the CPU cores of the GPU's NUMA node (NVML), so that the pinned host buffers allocated
    afterwards and the copy-issuing thread are local to the GPU's PCIe root: with eight ranks pumping ~100 MB per step
    each, remote pages halve the host-to-device rate. Best effort; returns the number of cores or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


class Env:
    """Process-wide state shared by the workloads of one run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from rspl_slam_b200 import capi
        self.torch, self.dist, self.capi, self.args = torch, dist, capi, args
        self.rank, self.local_rank, self.world = _dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.cpu_affinity = _bind_near_gpu(self.local_rank) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = capi.Context(device=self.local_rank)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.opt = capi.make_options()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.peak, self.peak_src = _peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t.tolist()]

    def timed_steps(self, solve, steps):
        """`steps` solves on the solver stream, an L2 flush before each, CUDA events around each; returns
        (device seconds, wall seconds of the region, clock summary)."""
        torch = self.torch
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        self.barrier()
        with ClockSampler(self.local_rank) as clk:
            t0 = time.perf_counter()
            with torch.cuda.stream(self.stream):
                for i in range(steps):
                    self.flush.zero_()  # evict L2 between timed steps (not timed)
                    starts[i].record(self.stream)
                    solve()
                    ends[i].record(self.stream)
            self.barrier()
            wall = time.perf_counter() - t0
        dev_s = float(sum(s.elapsed_time(e) for s, e in zip(starts, ends))) * 1e-3
        return dev_s, wall, clk.summary()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()
        self.ctx.close()


def _local_roofline(env, batch, stats, prof, prof_steps, key):
    from rspl_slam_b200.roofline import local_class_bytes
    per_kernel = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps} for k, v in prof.items() if v[1]}
    cls_bytes = local_class_bytes(batch, stats)  # bytes per step of the linearise / Schur / back-sub classes
    groups = {"linearize (kb_linearize + kb_pose_blocks)": (("linearize", "pose_blocks"), cls_bytes["linearize"]),
              "schur (kt_schur_tile | kb_schur_prep + kb_schur_reduce, + reduced solve)":
                  (("schur_tile", "schur_prep", "schur_reduce", "reduced_solve", "dense_assemble"), cls_bytes["schur"]),
              "backsub (kt_backsub_rc | kb_backsub)": (("backsub_update_eval",), cls_bytes["backsub"])}
    best = None
    for name, (classes, nbytes) in groups.items():
        ms = sum(prof[c][0] for c in classes if c in prof) / prof_steps
        nl = sum(prof[c][1] for c in classes if c in prof) / prof_steps
        if nl == 0:
            continue
        per_kernel[name] = {"ms_per_step": ms, "algorithmic_bytes_per_step": nbytes,
                            "achieved_GBs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / env.peak}
        if best is None or ms > best[1]:
            best = (name, ms, nl, nbytes)
    dom_name, dom_ms, dom_launches, alg_bytes = best
    launch_s = dom_ms * 1e-3 / max(dom_launches, 1)
    alg_per_launch = alg_bytes / max(dom_launches, 1)
    achieved = alg_per_launch / launch_s / 1e9
    return {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak,
            "traffic": _traffic(key), "peak_source": env.peak_src, "algorithmic_bytes_per_launch": alg_per_launch,
            "launch_ms": 1e3 * launch_s, "launches_per_step": dom_launches, "per_kernel": per_kernel,
            "note": "achieved = SURVEY 8(d) contract bytes (materialised-W formulation) of the slowest kernel class / its CUDA-event "
                    "time; the kernels recompute instead of materialising, so real DRAM traffic (`traffic`) is lower, see DESIGN.md"}


def bench_units(env, key, steps, warmup, with_cpu):
    """c2 / c2p / frame1 (frames) and c4 / c1 / c3 (local windows): independent units, this rank's shard."""
    from rspl_slam_b200 import synth
    args, ctx, capi, opt = env.args, env.ctx, env.capi, env.opt
    rank, world = env.rank, env.world
    is_frame = key in ("c2", "c2p", "frame1")
    if is_frame:
        n_units = 1 if key == "frame1" else args.frames
        nl = args.frame_lines if key == "c2" else 0
        batch = synth.make_frame_batch(2, n_units, first_instance=rank * n_units, n_points=C2_POINTS, n_lines=nl)
        if key == "frame1":  # a single call, as the reference makes them: one CTA per frame
            opt = capi.make_options(frame_latency_mode=1)
        upload, solve = ctx.frame_batch_upload, ctx.frame_batch_solve
        out = ctx.alloc_frame_result(batch, pinned=True)
        download = lambda: ctx.frame_batch_download(out)
        oneshot = lambda b: ctx.frame_batch(b, opt, out)
        if len(batch.cameras) == 1:
            # one camera and all ->inlier flags true are the ABI defaults: pass NULL instead of copying zeros / ones
            batch.mono_cam = batch.stereo_cam = None
            if batch.mono_inlier.all() and batch.stereo_inlier.all():
                batch.mono_inlier = batch.stereo_inlier = None
            if batch.mline_begin is not None:
                batch.mline_cam = batch.sline_cam = None
                if batch.mline_inlier.all() and batch.sline_inlier.all():
                    batch.mline_inlier = batch.sline_inlier = None
    else:
        n_units = args.windows if key == "c4" else 1
        cfg = {"c4": 4, "c1": 1, "c3": 3}[key]
        batch, _ = synth.make_local_batch(cfg, n_units, first_instance=rank * n_units, **_local_kw(key))
        upload, solve = ctx.local_batch_upload, ctx.local_batch_solve
        out = ctx.alloc_local_result(batch, pinned=True)
        download = lambda: ctx.local_batch_download(out)
        oneshot = lambda b: ctx.local_batch(b, opt, out)
    pinned = _pin_batch(batch, capi)

    # ---- resident-input timing
    upload(pinned)
    launches0 = ctx.launch_count
    warm = max(warmup, 3)
    for _ in range(warm):
        solve(opt)
    ctx.sync()
    download()
    stats = out.stats.copy()
    launches_per_step = (ctx.launch_count - launches0) // warm
    dev_s, wall, clocks = env.timed_steps(lambda: solve(opt), steps)
    edges_lin = int(stats["edges_linearized"].sum())
    edges_eval = int(stats["edges_evaluated"].sum())
    lm_iters = int(stats["iters"].sum())
    lm_trials = int(stats["trials"].sum())

    # ---- end to end through the C-ABI with pinned host buffers
    for _ in range(2):
        oneshot(pinned)
    env.barrier()
    e2e_steps = max(3, min(steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        oneshot(pinned)
    env.barrier()
    e2e_s = time.perf_counter() - t0

    dev_s_max, e2e_s_max, wall_max = env.reduce([dev_s, e2e_s, wall], "MAX")
    edges_lin_all, lm_iters_all, edges_eval_all, lm_trials_all = env.reduce([edges_lin, lm_iters, edges_eval, lm_trials], "SUM")

    # roofline of the dominant kernel class: algorithmic bytes / CUDA-event time of that class,
    # measured in a separate profiled pass (event pairs around every launch); rank 0's kernels
    prof_steps = max(1, min(steps, 5))
    ctx.set_profiling(True)
    for _ in range(prof_steps):
        solve(opt)
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    if rank != 0:
        return None
    if is_frame:
        cnt = {BYTES_POSE_ONLY_STEREO: int(batch.stereo_begin[-1]), BYTES_POSE_ONLY_MONO: int(batch.mono_begin[-1])}
        if batch.mline_begin is not None:
            cnt[BYTES_POSE_ONLY_STEREO_LINE] = int(batch.sline_begin[-1])
            cnt[BYTES_POSE_ONLY_MONO_LINE] = int(batch.mline_begin[-1])
        # evaluations are spread over the edge classes in proportion to their counts
        alg_bytes = edges_eval * sum(bts * n for bts, n in cnt.items()) / max(sum(cnt.values()), 1)
        dom_ms = prof["frame_opt"][0] / prof_steps
        dom_launches = prof["frame_opt"][1] / prof_steps
        launch_s = dom_ms * 1e-3 / max(dom_launches, 1)
        achieved = alg_bytes / max(dom_launches, 1) / launch_s / 1e9
        fp64_peak, fp64_src = _fp64_peak()
        roof = {"bound": "hbm", "kernel": "ba::frame_opt_kernel", "achieved": achieved, "peak": env.peak, "unit": "GB/s",
                "frac": achieved / env.peak, "traffic": _traffic(key), "peak_source": env.peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(dom_launches, 1), "launch_ms": 1e3 * launch_s,
                "launches_per_step": dom_launches,
                "fp64": {"peak_tflops": fp64_peak, "peak_source": fp64_src,
                         "note": "the kernel is FP64-issue bound (real DRAM traffic = `traffic`); its FP64 pipe utilisation is in the "
                                 "ncu summary under profiles/"},
                "note": "achieved = SURVEY 8(d) contract bytes (every LM pass streams the edge records) / CUDA-event time; the kernel "
                        "keeps the frame's edges on chip, so real DRAM traffic (`traffic`) is ~1 % of that, see DESIGN.md"}
    else:
        roof = _local_roofline(env, batch, stats, prof, prof_steps, key)
    res = {
        "metric": METRIC, "value": edges_lin_all * steps / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * dev_s_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(key, args), "l2": "flushed (256 MiB write) between timed steps; inputs resident in HBM",
                   "parallelism": f"dp{world} (independent units, no collective)",
                   "timing": "CUDA events on the solver stream, max over ranks"},
        "lm_iters_per_sec": lm_iters_all * steps / dev_s_max,
        "lm_trials_per_sec": lm_trials_all * steps / dev_s_max,
        "edges_evaluated_per_sec": edges_eval_all * steps / dev_s_max,
        "units_per_sec": n_units * world * steps / dev_s_max,
        "e2e": {"value": edges_lin_all * e2e_steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(pinned.h2d_bytes()),
                "d2h_bytes_per_step": int(out.d2h_bytes()), "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                "api": "rspl_ba_%s_batch (pinned host buffers in, pinned host buffers out)" % ("frame" if is_frame else "local")},
        "gpu_launches": int(launches_per_step * steps),
        "roofline": roof,
        "clocks": clocks,
        "wall_s_timed_region": wall_max,
    }
    if key in ("c1", "c3", "frame1"):
        res["shim_latency"] = shim_latency(key)
    if key == "c4" and world == 1:
        res["post_ba_line_refresh"] = post_ba_line_refresh(env, ctx, batch, solve, opt)
    if with_cpu:
        res["cpu_baseline"] = cpu_baseline(key, args)
    return res


def post_ba_line_refresh(env, ctx, batch, solve, opt, pts_per_line=12, reps=5):
    """The step that follows the path (Map::UppdateMapline for every optimised line) on the RESIDENT result
    of the batch just solved: rspl_ba_local_batch_update_maplines, only the point lists travel. Every line gets
    `pts_per_line` random map points of its own window (synthetic lists: the generator has no points-on-line relation,
    so few of them pass the 0.2 m gate; the work per reference is the same)."""
    try:
        capi = env.capi
        solve(opt)
        env.torch.cuda.synchronize(env.dev)
        nl = int(batch.line_begin[-1])
        rng = np.random.default_rng(20261018)
        win = np.repeat(np.searchsorted(batch.line_begin, np.arange(nl), side="right") - 1, pts_per_line)
        span = (batch.point_begin[win + 1] - batch.point_begin[win]).astype(np.int64)
        index = capi.pinned_empty((nl * pts_per_line,), np.int32)
        index[:] = batch.point_begin[win] + rng.integers(0, 1 << 40, nl * pts_per_line) % np.maximum(span, 1)
        begin = capi.pinned_empty((nl + 1,), np.int32)
        begin[:] = np.arange(nl + 1, dtype=np.int64) * pts_per_line
        out = (capi.pinned_empty((6, nl), np.float64), capi.pinned_empty((nl,), np.uint8))
        ctx.local_update_maplines(begin, index, out=out)
        t0 = time.perf_counter()
        for _ in range(reps):
            _, ok, cnt = ctx.local_update_maplines(begin, index, out=out)
        dt = (time.perf_counter() - t0) / reps
        return {"api": "rspl_ba_local_batch_update_maplines (lines and points resident from the solve; pinned lists in, pinned results out)",
                "lines": nl, "point_references": int(nl * pts_per_line), "lines_refreshed": int(cnt), "ms_per_call": 1e3 * dt,
                "h2d_bytes": int(4 * (nl + 1) + 4 * nl * pts_per_line), "d2h_bytes": int(49 * nl + 4)}
    except Exception as e:  # an extra; it must not take the bench line down
        return {"error": repr(e)[:300]}


def shim_latency(key, reps=20):
    """Latency of ONE call through the reference signatures (include/rspl_ba/g2o_optimization_shim.hpp): container
    flattening + H2D + solve + D2H + scatter, measured by tests/shim/shim_driver.cpp in a subprocess."""
    try:
        import tempfile
        from rspl_slam_b200 import synth
        sys.path.insert(0, os.path.join(ROOT, "tests", "shim"))
        from shim_dump import dump_problem
        tmp = tempfile.mkdtemp(prefix="rspl_shim_")
        exe = os.path.join(tmp, "shim_driver")
        env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
        subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT}/include", f"-I{ROOT}/tests/shim", f"{ROOT}/tests/shim/shim_driver.cpp",
                        "-o", exe, f"-L{ROOT}/rspl_slam_b200", "-lrspl_ba", f"-Wl,-rpath,{ROOT}/rspl_slam_b200"], check=True, env=env,
                       capture_output=True)
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        if key == "frame1":
            dump_problem(1, synth.make_frame_problem(synth.config_seed(2, 0), n_points=C2_POINTS), fin)
            entry = "FrameOptimization"
        else:
            cfg = {"c1": 1, "c3": 3}[key]
            dump_problem(0, synth.make_local_problem(synth.config_seed(cfg, 0), **_local_kw(key)), fin)
            entry = "LocalmapOptimization"
        r = subprocess.run([exe, fin, fout, str(reps)], check=True, capture_output=True, text=True, env=env, timeout=300)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        d["entry_point"] = entry + " (reference signature, host containers in / out)"
        return d
    except Exception as e:  # the latency leg must not take the bench line down
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------
# C5: ONE global problem, landmarks partitioned over the ranks (strong scaling; NCCL all-reduces of the
# pose blocks, the Schur complement pieces and a few scalars per LM trial)
# ------------------------------------------------------------------------------------------------
def _comm_init(env, ctx):
    torch, dist, capi = env.torch, env.dist, env.capi
    ident = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device=env.dev)
    if env.world > 1:
        if env.rank == 0:
            ident.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, src=0)
    ctx.comm_init(env.world, env.rank, bytes(ident.cpu().numpy().tobytes()) if env.world > 1 else None)


def global_parity_check(env):
    """N >= 2: a test-scale global problem solved over all ranks; poses must be bit-identical on every rank and agree
    with the un-sharded solve (one-rank communicator on rank 0) within the parity tolerances."""
    from rspl_slam_b200 import synth
    from rspl_slam_b200.geometry import quat_angle
    from rspl_slam_b200.problem import LocalBatch, shard_landmarks
    torch, dist = env.torch, env.dist
    full = synth.make_global_problem(synth.config_seed(5, 77), n_kf=48, n_points=12000, n_lines=1200, loops=1)
    shard = shard_landmarks(full, env.rank, env.world)
    res = env.ctx.global_ba(LocalBatch.from_problems([shard.problem]), env.opt)
    poses = torch.from_numpy(np.ascontiguousarray(res.pose_twc)).to(env.dev)
    gathered = [torch.empty_like(poses) for _ in range(env.world)]
    dist.all_gather(gathered, poses)
    out = None
    if env.rank == 0:
        bit_identical = all(bool(torch.equal(g.view(torch.int64), gathered[0].view(torch.int64))) for g in gathered)
        solo = env.capi.Context(device=env.local_rank)
        solo.comm_init(1, 0, None)
        ref = solo.global_ba(LocalBatch.from_problems([full]), env.opt)
        solo.comm_destroy()
        solo.close()
        dp = float(np.abs(res.pose_twc[:3] - ref.pose_twc[:3]).max())
        dr = float(max(quat_angle(res.pose_twc[3:, i], ref.pose_twc[3:, i]) for i in range(res.pose_twc.shape[1])))
        out = {"problem": "48 KF / 12k points / 1.2k lines (C5 generator)", "ranks": env.world,
               "poses_bit_identical_across_ranks": bool(bit_identical),
               "max_pose_diff_vs_one_rank_m": dp, "max_rot_diff_vs_one_rank_rad": dr,
               "within_tolerance": bool(dp < 1e-5 and dr < 1e-5),
               "iters_equal": bool(np.array_equal(res.stats["iters"], ref.stats["iters"]))}
    env.barrier()
    return out


def bench_global(env, steps, warmup, with_cpu):
    from rspl_slam_b200 import synth
    from rspl_slam_b200.problem import LocalBatch, shard_landmarks
    args, ctx, capi, opt = env.args, env.ctx, env.capi, env.opt
    rank, world = env.rank, env.world
    _comm_init(env, ctx)
    parity = global_parity_check(env) if world > 1 else None
    full = synth.make_global_problem(synth.config_seed(5, 0), n_kf=args.kf, n_points=args.points, n_lines=args.lines, loops=3,
                                     closure_every=args.closures)
    shard = shard_landmarks(full, rank, world)
    batch = LocalBatch.from_problems([shard.problem])
    pinned = _pin_batch(batch, capi)
    out = ctx.alloc_local_result(batch, pinned=True)
    total_edges = full.n_edges
    n_pts, n_lns = len(full.point_id), len(full.line_id)
    del full

    warm = max(warmup, 1)
    ctx.global_upload(pinned)
    launches0, coll0 = ctx.launch_count, ctx.collective_count()
    for _ in range(warm):
        ctx.global_solve(opt)
    ctx.sync()
    ctx.global_download(out)
    stats = out.stats.copy()
    launches_per_step = (ctx.launch_count - launches0) // warm
    coll_per_step = (ctx.collective_count() - coll0) // warm
    dev_s, wall, clocks = env.timed_steps(lambda: ctx.global_solve(opt), steps)
    # end to end: host buffers in, host buffers out
    ctx.global_ba(pinned, opt)
    env.barrier()
    e2e_steps = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.global_upload(pinned)
        ctx.global_solve(opt)
        ctx.global_download(out)
    env.barrier()
    e2e_s = time.perf_counter() - t0
    dev_s_max, e2e_s_max, wall_max = env.reduce([dev_s, e2e_s, wall], "MAX")
    h2d, d2h = env.reduce([float(pinned.h2d_bytes()), float(out.d2h_bytes())], "SUM")
    # the LM statistics are global (every rank holds the same numbers)
    edges_lin, edges_eval = int(stats["edges_linearized"][0]), int(stats["edges_evaluated"][0])
    lm_iters, lm_trials = int(stats["iters"][0].sum()), int(stats["trials"][0].sum())
    # (the profiled pass is collective too: every rank runs it, rank 0 reports)
    ctx.set_profiling(True)
    ctx.global_solve(opt)
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    res = None
    if rank == 0:
        # stats count the linearised edges of ALL ranks; this rank's kernels touched its shard's share of them
        stats_local = stats.copy()
        stats_local["edges_linearized"] = (stats["edges_linearized"] * (batch.n_edges / max(total_edges, 1))).astype(
            stats["edges_linearized"].dtype)
        roof = _local_roofline(env, batch, stats_local, prof, 1, "c5")
        roof["note"] = "rank 0's kernels on its landmark shard; contract bytes of SURVEY 8(d)"
        res = {
            "metric": METRIC, "value": edges_lin * steps / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * dev_s_max / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name("c5", args),
                       "l2": "flushed (256 MiB write) between timed steps; inputs resident in HBM",
                       "problem": f"{n_pts} points / {n_lns} lines / {total_edges} constraints",
                       "parallelism": f"landmark-sharded x{world}, poses replicated, {coll_per_step} all-reduces per solve from the "
                                      "library's own NCCL communicator (rspl_ba_comm_init)",
                       "timing": "CUDA events on the solver stream, max over ranks"},
            "lm_iters_per_sec": lm_iters * steps / dev_s_max, "lm_trials_per_sec": lm_trials * steps / dev_s_max,
            "edges_evaluated_per_sec": edges_eval * steps / dev_s_max, "units_per_sec": steps / dev_s_max,
            "e2e": {"value": edges_lin * e2e_steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                    "api": "rspl_ba_global_upload / _solve / _download (pinned host buffers, all ranks)"},
            "gpu_launches": int(launches_per_step * steps), "collectives_per_step": int(coll_per_step),
            "comm": {"library": "NCCL via rspl_ba_comm_init (dlopen)", "ranks": world, "all_reduces_per_solve": int(coll_per_step),
                     "ms_per_solve_rank0": prof.get("collectives", (0.0, 0))[0]},
            "roofline": roof, "clocks": clocks, "wall_s_timed_region": wall_max,
        }
        if parity is not None:
            res["parity_check"] = parity
        if with_cpu:
            res["cpu_baseline"] = cpu_baseline("c5", args)
    env.barrier()
    ctx.comm_destroy()
    return res


def sub_steps(key, args):
    """(steps, warmup) of a sub-workload of the default run, bounded so that the whole run ends within minutes."""
    if key == "c2p":
        return max(3, min(args.steps, 20)), 3
    if key == "c4":
        return max(2, min(args.steps, 5)), 3
    if key in ("c1", "c3", "frame1"):
        return max(5, min(args.steps, 30)), 3
    if key == "c5":
        return max(1, min(args.steps, 3)), 1
    if key in ("tri", "ends"):
        return max(2, min(args.steps, 5)), 2
    return args.steps, args.warmup


def _log(msg):
    if os.environ.get("RSPL_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', 0)} t={time.perf_counter():.1f}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    _log("start")
    env = Env(args)
    _log("env ready")
    multi = env.world > 1
    keys = (ALL_MULTI if multi else ALL_N1) if args.workload == "all" else [args.workload]
    if multi:
        keys = [k for k in keys if k not in ("c1", "c3", "frame1")] or keys  # single units do not shard (replicas only)
    results = {}
    for i, key in enumerate(keys):
        steps, warmup = (args.steps, args.warmup) if i == 0 else sub_steps(key, args)
        with_cpu = not multi  # the CPU baseline is reported on rank 0 at N = 1 only
        t0 = time.perf_counter()
        _log(f"workload {key}: steps={steps} warmup={warmup}")
        if key == "c5":
            results[key] = bench_global(env, steps, warmup, with_cpu)
        elif key == "tri":
            results[key] = bench_tri(env, steps, warmup, with_cpu)
        elif key == "ends":
            results[key] = bench_ends(env, steps, warmup, with_cpu)
        else:
            results[key] = bench_units(env, key, steps, warmup, with_cpu)
        _log(f"workload {key} done")
        if env.rank == 0 and results[key] is not None:
            results[key]["bench_wall_s"] = time.perf_counter() - t0
    if env.rank == 0:
        line = dict(results[keys[0]])
        if len(keys) > 1:
            line["workloads"] = {k: results[k] for k in keys[1:]}
        print(json.dumps(line), flush=True)
    env.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "c2", "c2p", "c4", "c1", "c3", "frame1", "c5", "tri", "ends"],
                    help="all: C2 headline + every other configuration as `workloads` sub-objects")
    ap.add_argument("--frame-lines", type=int, default=C2_LINES, help="lines per frame (c2)")
    ap.add_argument("--kf", type=int, default=2000, help="keyframes of the global problem (c5)")
    ap.add_argument("--points", type=int, default=1_000_000, help="points of the global problem (c5)")
    ap.add_argument("--lines", type=int, default=100_000, help="lines of the global problem (c5)")
    ap.add_argument("--closures", type=int, default=0,
                    help="c5: every N-th keyframe block also shares landmarks with the next lap (loop closures; 0 = none)")
    ap.add_argument("--frames", type=int, default=C2_FRAMES, help="frames per GPU (c2)")
    ap.add_argument("--windows", type=int, default=1024, help="windows per GPU (c4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
