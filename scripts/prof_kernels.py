"""Developer tool: one pass over every hot kernel of the VQ path at the BASELINE shapes, for ncu captures
(profiles/r02_*).  usage: python scripts/prof_kernels.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vq_seg_b200 as V  # noqa: E402
from vq_seg_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(5)
# ---- config 2: 8 x 256 x 64 x 64 map, K = 512: eval forward, train forward + backward, EMA statistics on the NCHW map
xs = [torch.relu(torch.randn(8, 256, 64, 64, generator=g, device=dev)) for _ in range(3)]
e = xs[0].permute(0, 2, 3, 1).reshape(-1, 256)[:512] + 0.05 * torch.randn(512, 256, generator=g, device=dev)
m = V.VectorQuantizer(dim=256, num_embeddings=512).to(dev)
m.codebook.embedding.weight.data.copy_(e)
m.eval()
with torch.no_grad():
    for x in xs:
        q, idx, loss, usage = m(x)
m.train()
for x in xs[:2]:
    xg = x.clone().requires_grad_(True)
    q, idx, loss, usage = m(xg)
    (q.sum() + loss.sum()).backward()
xv = xs[0].reshape(8, 256, 4096).permute(0, 2, 1)
for det in (False, True):
    counts, sums = ops.code_stats(xv, idx.reshape(8, 4096), 512, det)
# cosine lookup of the same map
mc = V.VectorQuantizer(dim=256, num_embeddings=512, distance="cosine").to(dev)
mc.codebook.embedding.weight.data.copy_(e)
mc.eval()
with torch.no_grad():
    mc(xs[1])
# ---- config 4 shape (a 256 k-row slice of it): packed rows, K = 1024, D = 512: streaming pair filter, rescoring,
# atomic and ordered statistics
n, d, k = 262144, 512, 1024
rows = torch.randn(1, n, d, generator=g, device=dev)
means = rows[0, :k].clone()
blob = ops.prepare_codebook(means)
for _ in range(2):
    idx4, counts4 = ops.assign(rows, means, blob, ops.ALGO_AUTO)
for det in (False, True):
    c4, s4 = ops.code_stats(rows, idx4, k, det)
# ---- config 5 shape (slice): K = 65536, D = 256
e5 = torch.randn(65536, 256, generator=g, device=dev)
x5 = torch.randn(1, 37888, 256, generator=g, device=dev)
i5, c5 = ops.assign(x5, e5, ops.prepare_codebook(e5), ops.ALGO_AUTO)
torch.cuda.synchronize()
print("done", usage.item(), int(counts4.sum()), int(c5.sum()))
