"""CPU restatement of VQ_SEG's vector-quantisation bottleneck  -- TEST INFRASTRUCTURE.

This file is the checker for the CUDA path.  It is NOT shipped and NOT measured as
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import it.

What it restates (all citations relative to /root/reference):
  vector_quantizer/vq_img.py:10-20   sample_vectors / batched_sample_vectors
  vector_quantizer/vq_img.py:22-27   batched_bincount
  vector_quantizer/vq_img.py:29-63   kmeans
  vector_quantizer/vq_img.py:65-130  CosinesimCodebook.forward
  vector_quantizer/vq_img.py:133-190 EuclideanCodebook.forward / _kmeans_init
  vector_quantizer/vq_img.py:193-244 VectorQuantizer.forward (STE, commitment loss)
  vector_quantizer/__init__.py:5-32  make_vq_module / Identity  (host logic only)

The arithmetic of the reference lives in torch (third-party, present in this image:
torch 2.11.0+cu128; the reference's own pin is the commented-out torch==1.13.1+cu116 in
requirements.txt:1).  `torch.cdist(p=2)` is ATen `_euclidean_dist`, restated in
`euclidean_dist` below as the augmented GEMM it really is; the restatement is asserted
bit-equal to `torch.cdist` and to the live reference module in tests/test_oracle.py.

PARITY PIN: the reference ships no tests, golden vectors or fixtures, so
nothing inside the reference pins this path.  The pin used here is "outputs of the live
reference module imported from /root/reference and run on CPU in the authoring
container", committed as tests/golden/*.pt by tests/golden/make_golden.py.

The restatement deliberately never builds the N x K int64 one-hot (vq_img.py:169-170): the
gather `E[idx]` is bit-identical to `one_hot(idx).float() @ E` (each output element is one
1.0*e product plus exact zeros), which tests/test_oracle.py asserts against the reference.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

__all__ = [
    "euclidean_dist", "assign_euclidean", "assign_cosine", "codebook_forward",
    "vq_forward", "vq_backward", "code_usage_from_counts", "kmeans", "sample_indices",
    "OracleVectorQuantizer", "oracle_make_vq_module", "OracleIdentity", "l2norm",
]


def l2norm(t: torch.Tensor) -> torch.Tensor:
    """vq_img.py:7-8"""
    return torch.nn.functional.normalize(t, p=2, dim=-1)


def euclidean_dist(x: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """`torch.cdist(x, e, p=2)` as called at vq_img.py:167 and :39.

    x: (B, N, D) (any strides), e: (K, D).  ATen `cdist_impl` first expands both operands to the
    common batch and makes them contiguous, then takes the GEMM path (`_euclidean_dist`) when
    N > 25 or K > 25 (probed):
        [-2x, |x|^2, 1] @ [e, 1, |e|^2]^T  -> clamp_min(0) -> sqrt
    with both norms computed as pow(2).sum(-1) on the contiguous copies and the product run as a
    batched matmul (the batching matters: MKL picks its K-blocking per call shape).
    For N <= 25 and K <= 25 ATen uses its vectorised exact-difference kernel, which this oracle
    does not restate: that branch calls torch.cdist itself (it is unreachable for any real
    codebook, K >= 26).
    """
    n, k = x.shape[-2], e.shape[-2]
    if n > 25 or k > 25:
        batch = x.shape[:-2]
        xc = x.contiguous()
        ec = e.expand(*batch, k, e.shape[-1]).contiguous()
        x_norm = xc.pow(2).sum(dim=-1, keepdim=True)
        e_norm = ec.pow(2).sum(dim=-1, keepdim=True)
        x_aug = torch.cat([xc.mul(-2), x_norm, torch.ones_like(x_norm)], dim=-1)
        e_aug = torch.cat([ec, torch.ones_like(e_norm), e_norm], dim=-1)
        return x_aug.matmul(e_aug.mT).clamp_min_(0).sqrt_()
    return torch.cdist(x, e, p=2)


def assign_euclidean(x_bnc: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """vq_img.py:167-168.  x_bnc: (B, HW, C) (any strides), e: (K, C).  -> (B, HW) int64.
    argmin returns the FIRST minimal index (lowest index wins ties)."""
    return torch.argmin(euclidean_dist(x_bnc, e), dim=-1)


def assign_cosine(x_bnc: torch.Tensor, e_unit: torch.Tensor) -> torch.Tensor:
    """vq_img.py:97,102-107.  x_bnc is l2-normalised here; e_unit must already be unit rows."""
    b, hw, c = x_bnc.shape
    flat = l2norm(x_bnc).contiguous().view(b * hw, c)
    sim = torch.einsum("n d, e d -> n e", flat, e_unit)
    return torch.argmax(sim.view(b, hw, -1), dim=-1)


def code_usage_from_counts(counts: torch.Tensor, num_embeddings: int) -> torch.Tensor:
    """vq_img.py:173-175: percent of UNUSED codes, 0-dim fp32."""
    zero_cnt = (counts == 0).sum()
    return 100 * (zero_cnt / num_embeddings)


def codebook_forward(x_bnc: torch.Tensor, e: torch.Tensor, distance: str = "euclidean"):
    """EuclideanCodebook.forward (vq_img.py:160-177) / CosinesimCodebook.forward (:93-113)
    without the k-means hook.  Returns (quantized (B,HW,C), idx (B,HW) int64, counts (K,) int64,
    code_usage 0-dim fp32)."""
    x_bnc = x_bnc.float()
    k = e.shape[0]
    if distance == "euclidean":
        idx = assign_euclidean(x_bnc, e)
    elif distance == "cosine":
        idx = assign_cosine(x_bnc, e)
    else:
        raise KeyError(distance)
    quantized = e[idx]                                    # == one_hot(idx).float() @ e, bit-exact
    counts = torch.bincount(idx.reshape(-1), minlength=k)
    return quantized, idx, counts, code_usage_from_counts(counts, k)


def vq_forward(x: torch.Tensor, e: torch.Tensor, training: bool, commitment_weight: float = 1.0,
               distance: str = "euclidean"):
    """VectorQuantizer.forward (vq_img.py:228-244) for an already-initialised codebook.

    x: (B, C, H, W) any float dtype.  Returns dict with
      quantize (B,C,H,W) fp32, embed_index (B,H,W) int64, loss (1,) fp32, code_usage () fp32,
      counts (K,) int64.
    For distance='cosine' the caller must pass the l2-normalised codebook (the reference
    renormalises its weights in place on every forward, vq_img.py:100).
    """
    x = x.to(torch.float32)
    b, c, h, w = x.shape
    x_bnc = x.reshape(b, c, h * w).permute(0, 2, 1)       # view, like rearrange 'b c h w -> b (h w) c'
    quantized, idx, counts, usage = codebook_forward(x_bnc, e, distance)
    loss = torch.zeros(1, dtype=torch.float32)
    if training:
        quantized = x_bnc + (quantized - x_bnc)           # STE value: two fp32 roundings (:236)
        if commitment_weight > 0:
            loss = loss + torch.nn.functional.mse_loss(quantized, x_bnc) * commitment_weight
    quantize = quantized.permute(0, 2, 1).reshape(b, c, h, w)
    return dict(quantize=quantize, embed_index=idx.view(b, h, w), loss=loss, code_usage=usage,
                counts=counts)


def vq_backward(x: torch.Tensor, quantize_ste: torch.Tensor, grad_quantize: Optional[torch.Tensor],
                grad_loss: Optional[torch.Tensor], commitment_weight: float) -> torch.Tensor:
    """Gradient of the training-mode forward w.r.t. x (vq_img.py:235-240).
    d quantize / dx = identity (STE);  d loss / dx = w * 2 (x - q_ste) / numel."""
    x = x.to(torch.float32)
    gx = torch.zeros_like(x) if grad_quantize is None else grad_quantize.to(torch.float32).clone()
    if grad_loss is not None and commitment_weight > 0:
        coef = grad_loss.reshape(()).to(torch.float32) * (2.0 * commitment_weight / x.numel())
        gx = gx + coef * (x - quantize_ste)
    return gx


def sample_indices(num_samples: int, num: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """vq_img.py:10-17: randperm(N)[:num] if N >= num else randint(0, N, (num,))."""
    if num_samples >= num:
        return torch.randperm(num_samples, generator=generator)[:num]
    return torch.randint(0, num_samples, (num,), generator=generator)


def kmeans(samples: torch.Tensor, num_clusters: int, num_iters: int,
           use_cosine_sim: bool = False, init_indices: Optional[torch.Tensor] = None,
           return_history: bool = False):
    """kmeans (vq_img.py:29-63) on samples (N, D) (the reference flattens (1,B,HW,C) to (1,N,C)).

    `init_indices` injects the initial row choice (the reference draws it with randperm on the
    sample's device, which is not reproducible across devices).
    Returns (means (K,D) fp32, bins (K,) int64) -- bins belong to the LAST assignment.
    """
    samples = samples.reshape(-1, samples.shape[-1]).contiguous().float()
    n, dim = samples.shape
    if init_indices is None:
        init_indices = sample_indices(n, num_clusters)
    means = samples[init_indices]
    bins = torch.zeros(num_clusters, dtype=torch.int64)
    history = []
    for _ in range(num_iters):
        if use_cosine_sim:
            dists = samples @ means.t()
        else:
            dists = -euclidean_dist(samples.unsqueeze(0), means)[0]
        buckets = torch.argmax(dists, dim=-1)
        bins = torch.zeros(num_clusters, dtype=torch.int64).scatter_add_(0, buckets, torch.ones_like(buckets))
        zero_mask = bins == 0
        bins_min_clamped = bins.masked_fill(zero_mask, 1)
        new_means = torch.zeros(num_clusters, dim, dtype=samples.dtype)
        new_means.scatter_add_(0, buckets.unsqueeze(1).expand(n, dim).contiguous(), samples)
        new_means = new_means / bins_min_clamped.unsqueeze(-1)
        if use_cosine_sim:
            new_means = l2norm(new_means)
        means = torch.where(zero_mask.unsqueeze(-1), means, new_means)
        if return_history:
            history.append((buckets.clone(), bins.clone(), means.clone()))
    if return_history:
        return means, bins, history
    return means, bins


def ema_update(counts: torch.Tensor, sums: torch.Tensor, cluster_size: torch.Tensor, embed_avg: torch.Tensor,
               decay: float, eps: float):
    """EMA codebook update -- OPT-IN EXTENSION WITHOUT A REFERENCE COUNTERPART (parity unpinned): the reference
    stores `decay` / `eps` (vq_img.py:150-151) and never reads them.  This restates the standard VQ-VAE rule
    (ema_inplace + laplace_smoothing of lucidrains/vector-quantize-pytorch, the code base the reference derives
    from) for the tests of vqseg_ema_update_f32.  Returns (cluster_size, embed_avg, weight)."""
    k = counts.numel()
    cluster_size = cluster_size * decay + (1 - decay) * counts.float()
    embed_avg = embed_avg * decay + (1 - decay) * sums
    n = cluster_size.sum()
    cs = (cluster_size + eps) / (n + k * eps) * n
    return cluster_size, embed_avg, embed_avg / cs.unsqueeze(1)


class OracleIdentity(torch.nn.Module):
    """vector_quantizer/__init__.py:27-32"""

    def __init__(self):
        super().__init__()
        self.embedding = torch.nn.Identity()

    def forward(self, x):
        return self.embedding(x), None, None, None


class OracleVectorQuantizer(torch.nn.Module):
    """Module-shaped restatement of VectorQuantizer (vq_img.py:193-244) incl. autograd, used as the
    CPU baseline in bench.py (`cpu_baseline.kind == "port"`) and as the autograd checker.
    Same constructor signature and state_dict keys as the reference."""

    def __init__(self, dim, num_embeddings, embedding_dim=None, decay=0.8, eps=1e-5, kmeans_init=False,
                 kmeans_iters=10, distance="euclidean", commitment_weight=1, num_codebook=1):
        super().__init__()
        embedding_dim = embedding_dim if embedding_dim is not None else dim
        if distance not in ("euclidean", "cosine"):
            raise KeyError(distance)
        self.num_embeddings, self.eps, self.commitment_weight = num_embeddings, eps, commitment_weight
        self.distance = distance
        cb = torch.nn.Module()
        cb.embedding = torch.nn.Embedding(num_embeddings, embedding_dim)
        cb.kmeans_init, cb.kmeans_iters, cb.initted = kmeans_init, kmeans_iters, False
        cb.num_codebook, cb.decay = num_codebook, decay
        cb.num_embeddings, cb.embedding_dim = num_embeddings, embedding_dim
        if not kmeans_init:
            cb.embedding.weight.data.uniform_(-1 / num_embeddings, 1 / num_embeddings)
            cb.initted = True
        self.codebook = cb
        self.kmeans_init_indices = None   # test hook: injected init rows
        self.faithful_ops = False         # True: run the reference's literal op sequence (one_hot + matmul, :169-170)

    def forward(self, x):
        x = x.to(torch.float32)
        b, c, h, w = x.shape
        x_bnc = x.reshape(b, c, h * w).permute(0, 2, 1)
        cb = self.codebook
        weight = cb.embedding.weight
        src = l2norm(x_bnc) if self.distance == "cosine" else x_bnc
        if cb.kmeans_init and self.training and not cb.initted:
            means, _ = kmeans(src.detach(), cb.num_embeddings, cb.kmeans_iters,
                              use_cosine_sim=self.distance == "cosine",
                              init_indices=self.kmeans_init_indices)
            weight.data.copy_(means)
            cb.initted = True
        if self.distance == "cosine":
            weight.data.copy_(l2norm(weight.data))
        with torch.no_grad():
            if self.faithful_ops and self.distance == "euclidean":
                idx = torch.argmin(torch.cdist(x_bnc, weight, p=2), dim=-1)
            else:
                idx = (assign_euclidean(x_bnc, weight) if self.distance == "euclidean"
                       else assign_cosine(x_bnc, weight))
        if self.faithful_ops:
            onehot = torch.nn.functional.one_hot(idx, num_classes=cb.num_embeddings)
            quantized = torch.matmul(onehot.float(), weight)
        else:
            quantized = weight[idx]
        counts = torch.bincount(idx.reshape(-1), minlength=cb.num_embeddings)
        usage = code_usage_from_counts(counts, cb.num_embeddings)
        loss = torch.tensor([0.0], device=x.device, requires_grad=self.training, dtype=torch.float32)
        if self.training:
            quantized = x_bnc + (quantized - x_bnc).detach()
            if self.commitment_weight > 0:
                loss = loss + torch.nn.functional.mse_loss(quantized.detach(), x_bnc) * self.commitment_weight
        quantize = quantized.permute(0, 2, 1).reshape(b, c, h, w)
        return quantize, idx.view(b, h, w), loss, usage


def oracle_make_vq_module(vq_cfg: dict, encoder_channels, depth):
    """vector_quantizer/__init__.py:5-25 with a plain dict in place of EasyDict."""
    import copy
    ne = vq_cfg["num_embeddings"]
    if isinstance(ne, int):
        return torch.nn.ModuleList([OracleVectorQuantizer(**vq_cfg, dim=encoder_channels[i + 1]) for i in range(depth)])
    if isinstance(ne, list):
        assert depth == len(ne), "depth and length of vq_cfg.num_embeddings must to be same number"
        cfg = copy.deepcopy(dict(vq_cfg))
        lst = []
        for i, n in enumerate(ne):
            cfg["num_embeddings"] = n
            if n == 0:
                lst.append(OracleIdentity())
            elif n > 0:
                lst.append(OracleVectorQuantizer(**cfg, dim=encoder_channels[i + 1]))
            else:
                raise ValueError(f"{n} is not available number of embeddings")
        return torch.nn.ModuleList(lst)
    raise TypeError(f"{type(ne)} is not available type")
