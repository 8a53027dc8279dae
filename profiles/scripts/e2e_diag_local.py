"""Wall time of the one-shot local call on the C4 batch (pinned buffers) and of upload / solve / download alone.
usage: python profiles/scripts/e2e_diag_local.py [windows]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rspl_slam_b200 import capi, synth  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = capi.Context(device=0)
opt = capi.make_options()
batch, _ = synth.make_local_batch(4, nw)
pinned = bench._pin_batch(batch, capi)
out = ctx.alloc_local_result(batch, pinned=True)


def wall(fn, reps=5):
    for _ in range(2):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / reps


res = {"windows": nw, "oneshot_ms": wall(lambda: ctx.local_batch(pinned, opt, out)),
       "upload_ms": wall(lambda: ctx.local_batch_upload(pinned))}
ctx.local_batch_upload(pinned)
res["solve_ms"] = wall(lambda: (ctx.local_batch_solve(opt), ctx.sync()))
res["download_ms"] = wall(lambda: ctx.local_batch_download(out))
res["h2d_bytes"], res["d2h_bytes"] = pinned.h2d_bytes(), out.d2h_bytes()
print(json.dumps(res))
