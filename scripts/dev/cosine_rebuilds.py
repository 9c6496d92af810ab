"""Developer check: the cosine codebook renormalises its weights in place EVERY forward (vq_img.py:100); how often
does that change their bits (each change makes the prologue's guard rebuild the prepared blob in one block)?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vq_seg_b200 as V
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.relu(torch.randn(8, 256, 64, 64, generator=g, device=dev))
m = V.VectorQuantizer(dim=256, num_embeddings=512, distance="cosine").to(dev)
m.codebook.embedding.weight.data.copy_(torch.randn(512, 256, generator=g, device=dev))
m.eval()
prev = None
for i in range(8):
    with torch.no_grad():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m(x); b.record(); torch.cuda.synchronize()
    w = m.codebook.embedding.weight.data.clone()
    changed = -1 if prev is None else int((w != prev).any(dim=1).sum())
    prev = w
    blob = m.codebook._blob
    print(f"forward {i}: {a.elapsed_time(b) * 1e3:7.1f} us, rows whose bits changed in this forward's renormalisation: {changed}, "
          f"blob rebuilds so far: {blob[96:100].view(torch.int32).item()}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        with torch.no_grad():
            m(x)
    torch.cuda.synchronize()
print("kernels of one cosine eval forward (device time, mean of 5):")
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0 and e.self_device_time_total > 0:
        print(f"   {e.key[:90]:90s} {e.self_device_time_total / 5:8.1f} us per forward ({e.count // 5} launches)")
        tot += e.self_device_time_total / 5
print(f"   total device time per forward: {tot:.1f} us")
m.enable_cuda_graphs()
for _ in range(3):
    with torch.no_grad():
        m(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    with torch.no_grad():
        m(x)
b.record(); torch.cuda.synchronize()
print(f"with enable_cuda_graphs(): {a.elapsed_time(b) * 1e3 / 20:.1f} us per cosine eval forward")
