"""ncu target (round 2): one batched local solve of N C1-shaped windows (default 128), or `frame N` for the pose-only kernel.
usage: ncu ... python profiles/scripts/r02_ncu_target.py [local|frame|c3] [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rspl_slam_b200 import capi, synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "local"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ctx = capi.Context(device=0)
opt = capi.make_options()
if kind == "frame":
    lines = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    b = synth.make_frame_batch(2, n, n_points=400, n_lines=lines)
    ctx.frame_batch_upload(b)
    for _ in range(2):
        ctx.frame_batch_solve(opt)
elif kind == "c3":
    b, _ = synth.make_local_batch(3, 1, n_kf=20, n_points=10000, n_lines=1000)
    ctx.local_batch_upload(b)
    for _ in range(2):
        ctx.local_batch_solve(opt)
else:
    b, _ = synth.make_local_batch(4, n)
    ctx.local_batch_upload(b)
    for _ in range(2):
        ctx.local_batch_solve(opt)
ctx.sync()
ctx.close()
