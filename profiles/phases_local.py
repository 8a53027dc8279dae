import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from rspl_slam_b200 import capi, synth
ctx = capi.Context(0)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 148
batch, _ = synth.make_local_batch(4, nw)
ctx.local_batch_upload(batch)
for _ in range(2):
    ctx.local_batch_solve(); ctx.sync()
t0 = time.perf_counter(); ctx.local_batch_solve(); ctx.sync(); dt = time.perf_counter() - t0
res = ctx.local_batch_download(ctx.alloc_local_result(batch))
ph = ctx.local_phase_cycles()
names = ["linearize", "pose_blocks", "schur_prep", "schur_reduce", "cholesky", "update_backsub_eval", "decision_restore", "other"]
print(f"{nw} windows: {dt*1e3:.2f} ms; iters {res.stats['iters'].sum(axis=0)[:2]/nw}, trials {res.stats['trials'].sum(axis=0)[:2]/nw}")
for n, c in zip(names, ph):
    print(f"  {n:22s} {100*c/ph.sum():5.1f}%  {c/nw/1.965e3:9.1f} us/window")
