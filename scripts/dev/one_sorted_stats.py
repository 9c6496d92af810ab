"""Developer tool: one unordered and one ordered code_stats call on 4 M x 512 packed rows, K = 1024, uniform clusters
(target of an ncu capture of stats_sorted_sum_kernel / stats_ordered_sum_rows_kernel / the sort kernels)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = 4_000_000, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
rows = torch.randn(1, n, d, generator=g, device=dev)
idx = torch.randint(0, k, (1, n), generator=g, device=dev)
for _ in range(2):
    ops.code_stats(rows, idx, k, False)
    ops.code_stats(rows, idx, k, True)
torch.cuda.synchronize()
print("ok")
