import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from rspl_slam_b200 import capi, synth
ctx = capi.Context(0)
batch, _ = synth.make_local_batch(4, 1024)
pinned = bench._pin_batch(batch, capi)
out = ctx.alloc_local_result(batch, pinned=True)
opt = capi.make_options()
for _ in range(2):
    ctx.local_batch(pinned, opt, out)
for rep in range(3):
    t0 = time.perf_counter(); ctx.local_batch_upload(pinned); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    ctx.local_batch_solve(opt); ctx.sync(); t3 = time.perf_counter()
    ctx.local_batch_download(out); t4 = time.perf_counter()
    print("upload call %.1f ms (+sync %.1f)  solve %.1f  download %.1f  total %.1f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t4-t0)*1e3))
