"""ncu / timing target (round 2): one global-BA solve of the C5 problem (or a shorter chain: argv[1] = keyframes) on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rspl_slam_b200 import capi, synth  # noqa: E402
from rspl_slam_b200.problem import LocalBatch  # noqa: E402

kf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
full = synth.make_global_problem(synth.config_seed(5, 0), n_kf=kf, n_points=kf * 500, n_lines=kf * 50, loops=3)
ctx = capi.Context(device=0)
ctx.comm_init(1, 0, None)
b = LocalBatch.from_problems([full])
ctx.global_upload(b)
opt = capi.make_options()
for _ in range(2):
    ctx.global_solve(opt)
ctx.sync()
ctx.comm_destroy()
ctx.close()
