"""Where the end-to-end time of the one-shot frame call goes: wall time per rspl_ba_frame_batch call on the C2 batch
with pinned buffers, and of upload / solve / download alone through the staged entry points. Measured (round 2):
one-shot 4.03 ms; upload 2.11 ms (104 MB at 49 GB/s), solve 2.80 ms, download 0.11 ms; the same one-shot call with the
kernel launches skipped in a diagnostic build (copy pipeline alone) 2.67 ms.
usage: python profiles/scripts/e2e_diag.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rspl_slam_b200 import capi, synth  # noqa: E402

ctx = capi.Context(device=0)
opt = capi.make_options()
batch = synth.make_frame_batch(2, 4096, n_points=400, n_lines=60)
batch.mono_cam = batch.stereo_cam = batch.mono_inlier = batch.stereo_inlier = None
batch.mline_cam = batch.sline_cam = batch.mline_inlier = batch.sline_inlier = None
pinned = bench._pin_batch(batch, capi)
out = ctx.alloc_frame_result(batch, pinned=True)


def wall(fn, reps=20):
    for _ in range(3):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / reps


res = {"oneshot_ms": wall(lambda: ctx.frame_batch(pinned, opt, out)),
       "upload_ms": wall(lambda: ctx.frame_batch_upload(pinned))}
ctx.frame_batch_upload(pinned)
res["solve_ms"] = wall(lambda: (ctx.frame_batch_solve(opt), ctx.sync()))
res["download_ms"] = wall(lambda: ctx.frame_batch_download(out))
res["h2d_bytes"] = pinned.h2d_bytes()
print(json.dumps(res))
