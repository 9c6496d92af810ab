"""Developer tool: one big cluster (a given share of 10 M rows) among uniform ones -- the chain rate of the big-cluster
ordered-sum kernel in isolation."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = 10_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
def run(idx, det):
    for _ in range(2): ops.code_stats(rows, idx, k, det)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.code_stats(rows, idx, k, det); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[1]
base = None
for share in (0.0, 0.05, 0.2, 0.5):
    idx = torch.randint(0, k, (1, n), generator=g, device=dev)
    idx[torch.rand(1, n, generator=g, device=dev) < share] = 7
    big = int((idx == 7).sum().item())
    t = run(idx, True)
    if base is None: base = t
    print(f"share {share:.2f}: big cluster {big} rows, ordered {t:.2f} ms, atomic {run(idx, False):.2f} ms"
          + (f" -> ~{big / max(t - base, 1e-3) / 1e3:.0f} rows/us on the big chain" if share else ""), flush=True)
