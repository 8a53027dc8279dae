"""Driver for ncu captures of the batched local-BA kernels: uploads N C1-shaped windows and runs one solve
with plain launches (profiling mode disables the CUDA graph). Usage: python profiles/prof_batched.py [windows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rspl_slam_b200 import capi, synth  # noqa: E402

ctx = capi.Context(0)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batch, _ = synth.make_local_batch(4, nw)
ctx.local_batch_upload(batch)
ctx.set_profiling(True)
ctx.local_batch_solve()
ctx.sync()
print("ok")
