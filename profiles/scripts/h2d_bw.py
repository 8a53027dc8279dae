"""Host-to-device copy bandwidth of the box for several copy sizes (pinned memory, one stream): what bounds the
end-to-end numbers of the one-shot calls. usage: python profiles/scripts/h2d_bw.py"""
import json
import torch

out = {}
for mb in (1, 4, 16, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    reps = max(4, 512 // mb)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out[f"h2d_{mb}MB_GBs"] = round(n * reps / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
    a.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out[f"d2h_{mb}MB_GBs"] = round(n * reps / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
print(json.dumps(out))
