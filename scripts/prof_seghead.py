"""Developer tool: a few forward + backward passes of the VQ segmentation head kernels at the reference's
decoder-output size (for `ncu -k regex:dist_map` captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
b, c, hw, k = 4, 32, 256 * 256, 3
e = torch.rand(k, c, generator=g, device=dev)
for _ in range(4):
    x = torch.relu(torch.randn(b, c, hw, generator=g, device=dev))
    dist, score, idx, counts = ops.dist_score_map(x.permute(0, 2, 1), e)
    gs = torch.randn_like(score)
    gx, ge = ops.dist_map_bwd(gs, dist, x.permute(0, 2, 1), e, score)
torch.cuda.synchronize()
print("ok", counts.tolist(), float(ge.abs().max()))
