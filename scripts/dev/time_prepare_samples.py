import sys, torch
sys.path.insert(0, '/root/repo')
from vq_seg_b200 import ops
dev = torch.device('cuda:0')
x = torch.randn(1, 2_000_000, 512, device=dev)
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s = ops.prepare_samples(x); b.record(); torch.cuda.synchronize()
    print("prepare_samples 2M x 512:", a.elapsed_time(b), "ms")
    del s
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    s = ops.prepare_samples(x); torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0: print(e.key[:80], e.device_time_total / 1e3, "ms")
