"""Developer tool: one assignment at a slice of the config-4 shape (1 M x 512, K = 1024) -- target of ncu captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
n, d, k = 1 << 20, 512, 1024
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(1, n, d, generator=g, device=dev)
e = x[0, torch.randperm(n, generator=g, device=dev)[:k]].contiguous()
blob = ops.prepare_codebook(e)
for _ in range(3): idx, counts = ops.assign(x, e, blob, ops.ALGO_AUTO)
torch.cuda.synchronize()
print("ok", int(counts.max()))
