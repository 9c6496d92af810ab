"""Developer tool: second ncu pass of round 2 -- the kernels that changed after profiles/r02_ncu_full.md was captured
(packed-row statistics of NCHW maps, block-per-pixel-group l2norm, split-D mode + short-list kernel) and the STE
backward.  usage: python scripts/prof_kernels2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vq_seg_b200 as V  # noqa: E402
from vq_seg_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.relu(torch.randn(8, 256, 64, 64, generator=g, device=dev))
e = x.permute(0, 2, 3, 1).reshape(-1, 256)[:512] + 0.05 * torch.randn(512, 256, generator=g, device=dev)
m = V.VectorQuantizer(dim=256, num_embeddings=512).to(dev)
m.codebook.embedding.weight.data.copy_(e)
m.train()
gq = torch.randn(8, 256, 64, 64, generator=g, device=dev)
for _ in range(2):
    xg = x.clone().requires_grad_(True)
    q, idx, loss, usage = m(xg)
    ((q * gq).sum() + loss.sum()).backward()            # dense grad_q: the flat STE backward kernel
xv = x.reshape(8, 256, 4096).permute(0, 2, 1)
for det in (False, True):
    ops.code_stats(xv, idx.reshape(8, 4096), 512, det)  # pack_rows + row kernels
ops.l2norm_rows(xv)
# split-D mode: the 1024- and 2048-channel layers of config 3
for (b, c, hw) in ((4, 1024, 1024), (4, 2048, 256)):
    f = torch.relu(torch.randn(b, c, hw, generator=g, device=dev)).permute(0, 2, 1)
    cb = (f.reshape(-1, c)[:512] + 0.05 * torch.randn(512, c, generator=g, device=dev)).contiguous()
    blob = ops.prepare_codebook(cb)
    for _ in range(2):
        ops.assign(f, cb, blob, ops.ALGO_AUTO)
torch.cuda.synchronize()
print("done")
