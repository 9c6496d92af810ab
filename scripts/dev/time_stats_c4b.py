"""Developer tool: ordered statistics at the config-4 shape -- wall (CUDA events), host enqueue time, result equality
between the per-range launches and the single launch (VQSEG_STATS_ONE=1)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
idx = torch.randint(0, k, (1, n), generator=g, device=dev)
def run(tag):
    for _ in range(2): out = ops.code_stats(rows, idx, k, True)
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t0 = time.perf_counter(); out = ops.code_stats(rows, idx, k, True); h = time.perf_counter() - t0; b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); hs.append(h * 1e3)
    print(f"{tag}: {sorted(ts)[2]:.2f} ms by events, host enqueue {sorted(hs)[2]:.2f} ms", flush=True)
    return out
os.environ.pop("VQSEG_STATS_ONE", None)
c0, s0 = run("one launch per 64 MB range")
os.environ["VQSEG_STATS_ONE"] = "1"
c1, s1 = run("single launch")
print("equal:", bool(torch.equal(c0, c1) and torch.equal(s0, s1)))
ta = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.code_stats(rows, idx, k, False); b.record(); torch.cuda.synchronize(); ta.append(a.elapsed_time(b))
print(f"atomic: {sorted(ta)[2]:.2f} ms")
