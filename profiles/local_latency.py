"""Single-call latency of LocalmapOptimization through the C-ABI for small batches, both device paths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rspl_slam_b200 import capi, synth
ctx = capi.Context(0)
for label, kw in (("C1 (10 KF / 3k pts / 300 lines)", dict()), ("C3 (20 KF / 10k pts / 1k lines)", dict(n_kf=20, n_points=10000, n_lines=1000))):
    for nw in (1, 8):
        batch, _ = synth.make_local_batch(1, nw, **kw)
        for path in ("persistent", "batched"):
            os.environ["RSPL_BA_LOCAL_PATH"] = path
            ctx.local_batch_upload(batch)
            ts = []
            for _ in range(4):
                t0 = time.perf_counter(); ctx.local_batch_solve(); ctx.sync(); ts.append(time.perf_counter() - t0)
            out = ctx.alloc_local_result(batch)
            t0 = time.perf_counter(); ctx.local_batch(batch, out=out); e2e = time.perf_counter() - t0
            print(f"{label:34s} windows={nw} path={path:10s} solve {1e3*min(ts[1:]):8.2f} ms   one-shot call {1e3*e2e:8.2f} ms")
            if path == "batched" and nw == 1:
                ctx.set_profiling(True); ctx.local_batch_solve(); prof = ctx.get_profile(); ctx.set_profiling(False)
                print("      per class (ms, launches):", {k: (round(v[0], 3), v[1]) for k, v in prof.items() if v[1]})
