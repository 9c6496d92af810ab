import sys, os, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
import train_step_c3 as T
import argparse
dev = torch.device('cuda:0')
torch.cuda.set_device(0)
args = argparse.Namespace(steps=2, warmup=2, per_gpu_batch=4, size=512)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    T.run_arm("b200", args, dev, 0, 1)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=15, max_name_column_width=60))
