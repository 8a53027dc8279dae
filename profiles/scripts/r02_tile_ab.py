"""A/B of the Schur paths on the GPU box (scratch driver, round 2): legacy (kb_schur_prep + kb_schur_reduce + Z in HBM)
against the tiled fused kernel at several tile sizes, for the C4 batch and single C1 / C3 windows.
usage: python profiles/scripts/r02_tile_ab.py [n_windows]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rspl_slam_b200 import capi, synth  # noqa: E402
from rspl_slam_b200.problem import LocalBatch  # noqa: E402


def run(ctx, batch, opt, label, reps=3, prof=True):
    ctx.local_batch_upload(batch)
    out = ctx.alloc_local_result(batch)
    for _ in range(2):
        ctx.local_batch_solve(opt)
    ctx.sync()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.local_batch_solve(opt)
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    ctx.local_batch_download(out)
    res = {"label": label, "ms": 1e3 * min(ts), "ms_all": [1e3 * t for t in ts]}
    if prof:
        ctx.set_profiling(True)
        ctx.local_batch_solve(opt)
        p = ctx.get_profile()
        ctx.set_profiling(False)
        res["classes"] = {k: round(v[0], 3) for k, v in p.items() if v[1]}
    print(json.dumps(res), flush=True)
    return out


def main():
    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ctx = capi.Context(device=0)
    opt = capi.make_options()
    t0 = time.time()
    batch, _ = synth.make_local_batch(4, nw)
    print("generated", nw, "windows in", round(time.time() - t0, 1), "s", flush=True)
    os.environ["RSPL_BA_SCHUR"] = "legacy"
    ref = run(ctx, batch, opt, f"c4x{nw} legacy")
    os.environ.pop("RSPL_BA_SCHUR")
    for q in (48 << 10, 100 << 10, 200 << 10):
        os.environ["RSPL_BA_TILE_Q"] = str(q)
        out = run(ctx, batch, opt, f"c4x{nw} tiled Q={q >> 10}K")
        dp = float(np.abs(out.pose_twc - ref.pose_twc).max())
        same = all(np.array_equal(getattr(out, a), getattr(ref, a)) for a in ("mp_inlier", "sp_inlier", "ml_inlier", "sl_inlier"))
        print(json.dumps({"vs_legacy_max_pose_diff": dp, "inlier_sets_equal": bool(same),
                          "iters_equal": bool(np.array_equal(out.stats["iters"], ref.stats["iters"]))}), flush=True)
    os.environ.pop("RSPL_BA_TILE_Q")
    # single windows: C1 and C3
    for name, kw in (("c1", dict(seed=synth.config_seed(1, 0))),
                     ("c3", dict(seed=synth.config_seed(3, 0), n_kf=20, n_points=10000, n_lines=1000))):
        seed = kw.pop("seed")
        p = synth.make_local_problem(seed, **kw)
        b1 = LocalBatch.from_problems([p])
        os.environ["RSPL_BA_SCHUR"] = "legacy"
        r1 = run(ctx, b1, opt, f"{name} legacy", reps=5, prof=False)
        os.environ.pop("RSPL_BA_SCHUR")
        o1 = run(ctx, b1, opt, f"{name} tiled", reps=5, prof=False)
        print(json.dumps({"window": name, "vs_legacy_max_pose_diff": float(np.abs(o1.pose_twc - r1.pose_twc).max())}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
