"""Generates tests/golden/golden_seghead_f64_v2.pt: the segmentation-head gradients of cases.SEGHEAD_CASES evaluated
in float64 (the head's mathematics, vq_segmentation_head.py:93-119 / 160-192 / 236-250, with the direct-difference
cdist), next to the error of the live reference's own float32 gradients (golden_seghead_v1.pt) against them.

    python tests/golden/make_golden_seghead_f64.py

Why: where a pixel coincides with a prototype the reference's composite backward (cat / matmul / clamp / sqrt) divides
by a distance that is pure cancellation noise (~1e-3) and then cancels two huge terms: its float32 prototype gradient is
1.1e-4 (relative to the largest entry) away from the float64 value on `sh_dup`.  No float32 implementation can agree
with THAT to 1e-5, so the GPU test bounds the kernel's error against float64 by the reference's own.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402


def f64_grads(name):
    build, distance = cases.SEGHEAD_CASES[name]
    x, e = build()
    b, c, h, w = x.shape
    xg = x.double().requires_grad_(True)
    eg = e.double().requires_grad_(True)
    xv = xg.reshape(b, c, h * w).permute(0, 2, 1)
    if distance == "cosine":
        wt = F.normalize(eg.detach(), dim=-1).requires_grad_(True)
        dist = torch.einsum("bnd,ed->bne", F.normalize(xv, dim=-1), wt)
        idx = dist.argmax(-1)
    else:
        wt = eg
        dist = torch.cdist(xv, eg, p=2, compute_mode="donot_use_mm_for_euclid_dist")
        idx = dist.argmin(-1)
    q = xv + (wt[idx] - xv).detach()
    loss = F.mse_loss(q.detach(), xv)
    score = dist.permute(0, 2, 1).reshape(b, -1, h, w)
    if distance == "euclidean":
        score = 1 - score / score.sum(1, keepdim=True)
    score = torch.softmax(score, 1)
    q = q.permute(0, 2, 1).reshape(b, c, h, w)
    g = torch.Generator().manual_seed(4242)
    gs = torch.randn(score.shape, generator=g).double()
    gq = torch.randn(q.shape, generator=g).double()
    ((score * gs).sum() + (q * gq).sum() + 1.5 * loss).backward()
    return xg.grad, wt.grad


def main():
    ref = torch.load(os.path.join(HERE, "golden_seghead_v1.pt"), weights_only=False)["seghead"]
    out = {}
    for name in cases.SEGHEAD_CASES:
        gx, gw = f64_grads(name)
        rec = ref[name]
        out[name] = {"gx_f64": gx, "gw_f64": gw,
                     "ref_gw_err": ((rec["gw"].double() - gw).abs().max() / gw.abs().max()).item(),
                     "ref_gx_err": ((rec["gx"].double() - gx).abs().max() / gx.abs().max()).item()}
        print(f"{name:12s} reference fp32 vs fp64: gw {out[name]['ref_gw_err']:.3e}  gx {out[name]['ref_gx_err']:.3e}")
    path = os.path.join(HERE, "golden_seghead_f64_v2.pt")
    torch.save({"seghead_f64": out}, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
