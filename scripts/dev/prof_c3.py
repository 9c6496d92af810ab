"""Developer tool: torch profiler over a few training steps of one arm of scripts/train_step_c3.py."""
import sys, os, argparse, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import train_step_c3 as T
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0'); torch.cuda.set_device(0)
arm = sys.argv[1] if len(sys.argv) > 1 else "b200_noema"
args = argparse.Namespace(steps=3, warmup=3, per_gpu_batch=4, size=512)
models, opts, step = T.build(arm, args, dev, 0, 1)
for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
ka = prof.key_averages()
print(f"arm {arm}: wall {wall * 1e3 / 3:.1f} ms per step under the profiler")
rows = [(e.key, e.count, e.device_time_total / 1e3, e.cpu_time_total / 1e3) for e in ka]
tot_dev = sum(e.self_device_time_total for e in ka) / 1e3
print(f"total device time {tot_dev / 3:.1f} ms per step")
print("--- ours (vqseg / custom functions)")
for k, c, d, cpu in sorted(rows, key=lambda r: -r[2]):
    if "vqseg" in k or "Fused" in k or "StraightThrough" in k or "EvalGather" in k:
        print(f"{k[:90]:90s} calls {c:5d}  device {d / 3:8.3f} ms/step  cpu {cpu / 3:8.3f} ms/step")
print("--- top device time")
for k, c, d, cpu in sorted(rows, key=lambda r: -r[2])[:14]:
    print(f"{k[:90]:90s} calls {c:5d}  device {d / 3:8.3f} ms/step  cpu {cpu / 3:8.3f} ms/step")
print("--- top cpu time")
for k, c, d, cpu in sorted(rows, key=lambda r: -r[3])[:14]:
    print(f"{k[:90]:90s} calls {c:5d}  device {d / 3:8.3f} ms/step  cpu {cpu / 3:8.3f} ms/step")
