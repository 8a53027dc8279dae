"""The library compiled with -DRSPL_BA_CHECKED (bounds / invariant asserts inside the kernels, `BA_CHECK` in
ba_math.cuh) runs the frame, local, large-window, global, triangulation and line-endpoint paths on small problems without tripping an assert and
with the same results as the shipped build. Substitute for compute-sanitizer memcheck, which is closed on the pool."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
from rspl_slam_b200 import capi, synth
from rspl_slam_b200.problem import FrameBatch, LocalBatch
ctx = capi.Context(device=0)
out = {}
fb = synth.make_frame_batch(2, 24, n_points=150, stereo_frac=0.8, n_lines=12)
r = ctx.frame_batch(fb)
out["frame_pose"] = r.pose_twc.tolist()
out["frame_inl"] = int(r.stereo_inlier.sum()) + int(r.mono_inlier.sum())
lat = capi.make_options(); lat.frame_latency_mode = 1
r1 = ctx.frame_batch(FrameBatch.from_problems([synth.make_frame_problem(synth.config_seed(2, 5), n_points=200)]), lat)
out["frame1_pose"] = r1.pose_twc.tolist()
probs = [synth.make_local_problem(synth.config_seed(1, 300 + i), n_kf=4 + i, n_points=120 + 40 * i, n_lines=15 + 5 * i) for i in range(3)]
lr = ctx.local_batch(LocalBatch.from_problems(probs))
out["local_pose"] = lr.pose_twc.tolist()
lb = LocalBatch.from_problems(probs)
nl, npnt = int(lb.line_begin[-1]), int(lb.point_begin[-1])
rng = np.random.default_rng(3)
pb = np.arange(nl + 1, dtype=np.int32) * 6
pw = np.repeat(np.searchsorted(lb.line_begin, np.arange(nl), side="right") - 1, 6)
pi = (lb.point_begin[pw] + rng.integers(0, 1 << 30, nl * 6) %% (lb.point_begin[pw + 1] - lb.point_begin[pw])).astype(np.int32)
re_, rok, _ = ctx.local_update_maplines(pb, pi)
out["resident_ends"] = [re_.tolist(), int(rok.sum())]
out["local_inl"] = int(lr.sp_inlier.sum()) + int(lr.mp_inlier.sum()) + int(lr.sl_inlier.sum()) + int(lr.ml_inlier.sum())
big = synth.make_global_problem(synth.config_seed(5, 1), n_kf=80, n_points=6000, n_lines=600)
gr = ctx.local_batch(LocalBatch.from_problems([big]))
out["big_pose"] = gr.pose_twc.tolist()
tb = synth.make_triangulation_batch(9, n_points=3000)
tx, tok, _ = ctx.triangulate_points(tb["obs_begin"], tb["obs_frame"], tb["obs_uv"], tb["frame_twc"], tb["cam5"])
out["tri"] = [tx.tolist(), int(tok.sum())]
mb = synth.make_mapline_batch(10, n_lines=3000)
me, mok, _ = ctx.update_maplines(mb["line_wd"], mb["pt_begin"], mb["pt_index"], mb["point_xyz"])
out["ends"] = [me.tolist(), int(mok.sum())]
ctx.close()
print("RESULT" + json.dumps(out))
"""


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["RSPL_BA_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "child failed (a BA_CHECK assert traps the kernel):\n" + r.stdout[-3000:] + r.stderr[-3000:]
    assert "BA_CHECK failed" not in r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


@pytest.mark.gpu
def test_checked_build_runs_clean_and_matches_shipped_build():
    from rspl_slam_b200 import build
    lib = build.build_checked()
    checked = _run(lib)
    shipped = _run(None)
    assert checked == shipped  # same kernels, same bits: the asserts only observe
