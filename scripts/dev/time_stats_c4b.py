"""Developer tool: ordered (bit-exact) against atomic per-code statistics at the config-4 shape (10 M packed rows,
D = 512, K = 1024) for uniform and skewed cluster sizes: wall by CUDA events and host enqueue time."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
def run(tag, idx, det):
    for _ in range(2): ops.code_stats(rows, idx, k, det)
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t0 = time.perf_counter(); ops.code_stats(rows, idx, k, det); h = time.perf_counter() - t0; b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); hs.append(h * 1e3)
    print(f"{tag}: {sorted(ts)[2]:.2f} ms by events, host enqueue {sorted(hs)[2]:.2f} ms", flush=True)
for name, p in (("uniform", 1.0), ("skewed u^2", 2.0), ("skewed u^4", 4.0)):
    idx = (torch.rand(1, n, generator=g, device=dev) ** p * k).long().clamp_(0, k - 1)
    big = int(torch.bincount(idx.reshape(-1), minlength=k).max().item())
    print(f"--- {name}: largest cluster {big} rows ({big * k / n:.1f}x the mean)")
    run("  ordered", idx, True)
    run("  atomic ", idx, False)
