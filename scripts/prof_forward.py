"""Developer tool: a few fused forwards of the C2 batch (for `ncu -k regex:...` captures).
usage: python scripts/prof_forward.py [eval|train]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
mode = ops.MODE_TRAIN if (len(sys.argv) > 1 and sys.argv[1] == "train") else ops.MODE_EVAL
g = torch.Generator(device="cuda").manual_seed(5)
xs = [torch.relu(torch.randn(8, 256, 4096, generator=g, device=dev)) for _ in range(6)]
e = xs[0].permute(0, 2, 1).reshape(-1, 256)[:512] + 0.05 * torch.randn(512, 256, generator=g, device=dev)
blob = ops.prepare_codebook(e)
for x in xs:
    q, idx, mse, usage = ops.vq_forward(x.permute(0, 2, 1), e, blob, mode, ops.ALGO_AUTO)
torch.cuda.synchronize()
print("usage", usage.item(), "flag-free check idx sum", idx.sum().item())
