"""Developer tool: one ordered code_stats call with a 20 % cluster (target of an ncu capture of the big-cluster kernel)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = 4_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
idx = torch.randint(0, k, (1, n), generator=g, device=dev)
idx[torch.rand(1, n, generator=g, device=dev) < 0.2] = 7
for _ in range(2): ops.code_stats(rows, idx, k, True)
torch.cuda.synchronize()
print("ok")
