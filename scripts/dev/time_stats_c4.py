"""Developer tool: per-kernel device times of the ordered (bit-exact) and atomic per-code statistics at the config-4
shape (10 M packed rows, D = 512, K = 1024), via the torch profiler."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
idx = torch.randint(0, k, (1, n), generator=g, device=dev)
for det in (False, True):
    for _ in range(2): ops.code_stats(rows, idx, k, det)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): ops.code_stats(rows, idx, k, det)
        torch.cuda.synchronize()
    print("deterministic" if det else "atomic", "(device time per call, mean of 3):")
    tot = 0.0
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_time_total > 0:
            print(f"   {e.key[:90]:90s} {e.device_time_total / 3:10.1f} us  ({e.count // 3} launches)")
            tot += e.device_time_total / 3
    print(f"   total {tot:.1f} us = {4.0 * n * d / tot / 1e6:.2f} TB/s of row data")
