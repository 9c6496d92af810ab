"""Developer tool: CUDA-event timings of the statistics and l2norm kernels at the config-2 / config-4 shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[n // 2]
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.relu(torch.randn(8, 256, 4096, generator=g, device=dev)); xv = x.permute(0, 2, 1)
idx = torch.randint(0, 512, (8, 4096), generator=g, device=dev)
print("C2 map (NCHW 8x256x64x64, K=512): code_stats atomic %.1f us, ordered %.1f us, l2norm %.1f us" % (
    t(lambda: ops.code_stats(xv, idx, 512, False)), t(lambda: ops.code_stats(xv, idx, 512, True)), t(lambda: ops.l2norm_rows(xv))))
c1, s1 = ops.code_stats(xv, idx, 512, True); c2, s2 = ops.code_stats(xv.contiguous(), idx, 512, True)
print("   ordered sums on the NCHW view == on packed rows:", bool(torch.equal(s1, s2) and torch.equal(c1, c2)))
for (b, c, hw) in [(4, 512, 4096), (4, 1024, 1024), (4, 2048, 256)]:
    x = torch.relu(torch.randn(b, c, hw, generator=g, device=dev)); xv = x.permute(0, 2, 1)
    idx = torch.randint(0, 512, (b, hw), generator=g, device=dev)
    print(f"C3 layer {b}x{c}x{hw}: code_stats atomic %.1f us, ordered %.1f us, l2norm %.1f us" % (
        t(lambda: ops.code_stats(xv, idx, 512, False)), t(lambda: ops.code_stats(xv, idx, 512, True)), t(lambda: ops.l2norm_rows(xv))))
n, d, k = 1 << 20, 512, 1024
rows = torch.randn(1, n, d, generator=g, device=dev); idx = torch.randint(0, k, (1, n), generator=g, device=dev)
ta, to = t(lambda: ops.code_stats(rows, idx, k, False), 5), t(lambda: ops.code_stats(rows, idx, k, True), 5)
print(f"C4 slice (1M packed rows, D=512, K=1024): atomic {ta:.0f} us ({4.0 * n * d / ta / 1e6:.2f} TB/s), ordered {to:.0f} us ({4.0 * n * d / to / 1e6:.2f} TB/s)")
# kernel-only times (torch profiler, device side)
from torch.profiler import profile, ProfilerActivity
x = torch.relu(torch.randn(8, 256, 4096, generator=g, device=dev)); xv = x.permute(0, 2, 1)
idx = torch.randint(0, 512, (8, 4096), generator=g, device=dev)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.code_stats(xv, idx, 512, False); ops.code_stats(xv, idx, 512, True); ops.l2norm_rows(xv)
    torch.cuda.synchronize()
print("kernel times on the C2 map (device, mean of 5):")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"   {e.key[:100]:100s} {e.device_time_total / e.count:8.1f} us x {e.count // 5} per call set")
