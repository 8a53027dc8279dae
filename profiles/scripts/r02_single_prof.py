"""Per-kernel-class times of single-window solves (C1, C3) on the GPU box: host-driven path with event pairs around
every launch (profiling switches the CUDA graph off), and the graph path's end-to-end solve time beside it."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rspl_slam_b200 import capi, synth  # noqa: E402
from rspl_slam_b200.problem import LocalBatch  # noqa: E402

ctx = capi.Context(device=0)
opt = capi.make_options()
for name, cfg, kw in (("c1", 1, {}), ("c3", 3, dict(n_kf=20, n_points=10000, n_lines=1000))):
    b = LocalBatch.from_problems([synth.make_local_problem(synth.config_seed(cfg, 0), **kw)])
    ctx.local_batch_upload(b)
    out = ctx.alloc_local_result(b)
    for mode in ("graph", "host"):
        if mode == "host":
            os.environ["RSPL_BA_GRAPH"] = "off"
        else:
            os.environ.pop("RSPL_BA_GRAPH", None)
        for _ in range(3):
            ctx.local_batch_solve(opt)
        ctx.sync()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            ctx.local_batch_solve(opt)
            ctx.sync()
            ts.append(time.perf_counter() - t0)
        ctx.local_batch_download(out)
        print(json.dumps({"window": name, "mode": mode, "ms_min": 1e3 * min(ts), "ms_med": 1e3 * sorted(ts)[5],
                          "iters": out.stats["iters"][0].tolist(), "trials": out.stats["trials"][0].tolist()}), flush=True)
    ctx.set_profiling(True)
    ctx.local_batch_solve(opt)
    p = ctx.get_profile()
    ctx.set_profiling(False)
    print(json.dumps({"window": name, "classes_ms": {k: round(v[0], 3) for k, v in p.items() if v[1]},
                      "launches": {k: v[1] for k, v in p.items() if v[1]}}), flush=True)
ctx.close()
