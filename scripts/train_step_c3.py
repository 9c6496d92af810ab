"""BASELINE config 3: the reference's semi-supervised training step with the VQ bottleneck swapped -- "train img/s".

    python scripts/train_step_c3.py [--steps 10] [--warmup 3] [--per-gpu-batch 4] [--size 512] [--arms b200,torch,identity]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_step_c3.py ...

What is re-driven is the step body of /root/reference/train_vqreptunet1x1v2.py:129-202 (cross pseudo supervision):
two models; under no_grad both label the unlabelled batch in eval mode; under fp16 autocast both run the labelled and
the unlabelled batch in train mode (4 train-mode + 2 eval-mode forwards); CPS loss with confidence masking
(score_mask, :43-46), supervised 0.5 CE + dice, the commitment losses of the VQ layers; GradScaler; two Adam steps.
The model follows VQRePTUnet1x1v2.forward (models/networks/modified_vqunet/net.py:217-247): encoder features l3-l5
(512 / 1024 / 2048 channels at 1/8, 1/16, 1/32 resolution) each go through their own VectorQuantizer
(num_embeddings [0, 0, 512, 512, 512], config/vqreptunet1x1v2.json), the quantized maps feed a U-Net decoder and a
1x1 segmentation head.

/root/reference does not exist on the GPU box, so the ENCODER / DECODER here are stand-ins of the same layer shapes
in stock PyTorch (torchvision's resnet50 without weights + a plain U-Net decoder); they are the callers, not the
product (out of scope), and tests/test_install_reference.py builds the real VQRePTUnet1x1v2 with these
codebooks where the reference is present.  The prototype loss (ReliablePrototypeLossv2) is not on the VQ path and is
left out.  Data: synthetic rand images / randint labels (encoder_weights=None), 4 + 4 images per GPU like the config.

Arms (same step, same seeds):
  b200     vq_seg_b200.VectorQuantizer through make_vq_module (kmeans_init on the first training forward, the opt-in
           EMA codebook update; under torchrun both use the NCCL all-reduce of per-code counts and sums)
  torch    the reference's literal op sequence (cdist, argmin, one_hot, matmul, bincount, mse_loss) in torch eager
           on the same GPU: the bar on the box
  identity no quantisation: what the step costs without the VQ layers
One step processes 2 x per_gpu_batch x world images (labelled + unlabelled); "train img/s" counts the labelled
ones like SURVEY 8d (32 / t_step at 8 GPUs).
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# ---- stand-in callers (stock PyTorch) --------------------------------------------------------------------------------
class Encoder(nn.Module):
    """resnet50 feature pyramid: channels (3, 64, 256, 512, 1024, 2048) at strides (1, 2, 4, 8, 16, 32)."""

    def __init__(self):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu)
        self.pool = r.maxpool
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])

    def out_channels(self):
        return (3, 64, 256, 512, 1024, 2048)

    def forward(self, x):
        feats = [x]
        x = self.stem(x); feats.append(x)
        x = self.pool(x)
        for layer in self.layers:
            x = layer(x); feats.append(x)
        return feats


class DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin + cskip, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                  nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv(x)


class VQUnet(nn.Module):
    """the shape of VQRePTUnet1x1v2 (net.py:180-247) around `codebook`, an nn.ModuleList of five VQ modules"""

    def __init__(self, codebook, num_classes=3):
        super().__init__()
        self.encoder = Encoder()
        ch = self.encoder.out_channels()
        self.codebook = codebook
        dec = [c // 2 for c in ch[1:]][::-1]                      # (1024, 512, 256, 128, 32), net.py:207-209
        skips = list(ch[1:-1][::-1]) + [0]                         # (1024, 512, 256, 64, 0)
        cins = [ch[-1]] + dec[:-1]
        self.blocks = nn.ModuleList([DecoderBlock(ci, cs, co) for ci, cs, co in zip(cins, skips, dec)])
        self.segmentation_head = nn.Conv2d(dec[-1], num_classes, 1, bias=False)

    def forward(self, x):
        features = self.encoder(x)[1:]
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)
        usage = []
        for i in range(len(features)):
            quantize, _idx, commitment_loss, code_usage = self.codebook[i](features[i])
            features[i] = quantize
            if commitment_loss is not None:
                loss = loss + commitment_loss
            if code_usage is not None:
                usage.append(code_usage.detach())
        loss = loss / len(features)
        feats = features[::-1]
        y = feats[0].to(feats[1].dtype) if feats[0].dtype != feats[1].dtype else feats[0]
        for i, blk in enumerate(self.blocks):
            skip = feats[i + 1] if i + 1 < len(feats) else None
            y = blk(y, skip.to(y.dtype) if skip is not None else None)
        return self.segmentation_head(y), loss, usage


class TorchVQ(nn.Module):
    """the reference's op sequence (vq_img.py:160-177, 228-244) in torch eager on the GPU -- the comparison arm"""

    def __init__(self, dim, num_embeddings, commitment_weight=1.0):
        super().__init__()
        self.embedding = nn.Embedding(num_embeddings, dim)
        self.embedding.weight.data.uniform_(-1 / num_embeddings, 1 / num_embeddings)
        self.num_embeddings, self.commitment_weight = num_embeddings, commitment_weight

    def forward(self, x):
        b, c, h, w = x.shape
        flat = x.float().permute(0, 2, 3, 1).reshape(b, h * w, c)
        dist = torch.cdist(flat, self.embedding.weight, p=2)
        idx = torch.argmin(dist, dim=-1)
        quantize = torch.matmul(F.one_hot(idx, num_classes=self.num_embeddings).float(), self.embedding.weight)
        counts = torch.bincount(idx.reshape(-1), minlength=self.num_embeddings)
        usage = 100 * ((counts == 0).sum() / self.num_embeddings)
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)
        if self.training:
            quantize = flat + (quantize - flat).detach()
            loss = loss + F.mse_loss(quantize.detach(), flat) * self.commitment_weight
        return quantize.reshape(b, h, w, c).permute(0, 3, 1, 2), idx.reshape(b, h, w), loss, usage


class NoVQ(nn.Module):
    def forward(self, x):
        return x, None, None, None


def make_codebooks(arm, dev, world):
    ch, depth, ks = (3, 64, 256, 512, 1024, 2048), 5, [0, 0, 512, 512, 512]
    if arm.startswith("b200"):
        import vq_seg_b200 as V
        from vq_seg_b200 import distributed as VD
        cb = V.make_vq_module({"num_embeddings": ks, "distance": "euclidean", "kmeans_init": True}, ch, depth)
        for m in cb:
            if isinstance(m, V.VectorQuantizer):
                m.codebook.embedding.weight.requires_grad_(False)     # detached in training (vq_img.py:236-239): no gradient
                if world > 1:
                    m.codebook.kmeans_reduce_fn = VD.allreduce_code_stats
                if arm != "b200_noema":                               # (b200_noema: the reference's semantics, no codebook update)
                    m.enable_ema(reduce_fn=VD.allreduce_code_stats if world > 1 else None)
        return cb.to(dev)
    if arm == "torch":
        mods = [NoVQ() if k == 0 else TorchVQ(c, k) for c, k in zip(ch[1:], ks)]
        for m in mods:
            if isinstance(m, TorchVQ):
                m.embedding.weight.requires_grad_(False)
        return nn.ModuleList(mods).to(dev)
    return nn.ModuleList([NoVQ() for _ in ks]).to(dev)


def score_mask(pred, pseudo, th):
    """train_vqreptunet1x1v2.py:43-46: pseudo labels whose softmax confidence is below th become ignore (255)"""
    conf = torch.softmax(pred.float(), dim=1).max(dim=1)[0]
    return torch.where(conf < th, torch.full_like(pseudo, 255), pseudo)


def dice_loss(pred, target, num_classes=3, eps=1e-6):
    valid = (target != 255)
    t = F.one_hot(torch.where(valid, target, torch.zeros_like(target)), num_classes).permute(0, 3, 1, 2).float() * valid.unsqueeze(1)
    p = torch.softmax(pred.float(), dim=1) * valid.unsqueeze(1)
    inter = (p * t).sum(dim=(0, 2, 3))
    return 1 - ((2 * inter + eps) / (p.sum(dim=(0, 2, 3)) + t.sum(dim=(0, 2, 3)) + eps)).mean()


def build(arm, args, dev, rank, world):
    """(models, step): the two CPS models of one arm and the closure that runs one training step"""
    torch.manual_seed(1234)                                   # same initial weights in every arm and on every rank
    models = [VQUnet(make_codebooks(arm, dev, world)).to(dev) for _ in range(2)]
    if world > 1:
        # (two forwards of each model precede the backward: DDP's per-forward buffer broadcast would rewrite BatchNorm
        # statistics the first forward's graph still needs; the EMA buffers stay equal through the all-reduced statistics)
        models = [nn.parallel.DistributedDataParallel(m, device_ids=[dev.index], broadcast_buffers=False) for m in models]
    opts = [torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4) for m in models]
    # (the default initial scale 2^16 overflows fp16 on this untrained net for the first ~6 steps: skipped optimizer
    # steps would make the short timed window incomparable between arms)
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 10)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    nb, s = args.per_gpu_batch, args.size
    ce = nn.CrossEntropyLoss(ignore_index=255)
    nb_ = nb

    def step():
        l_input = torch.rand(nb, 3, s, s, device=dev, generator=g)
        l_target = torch.randint(0, 3, (nb, s, s), device=dev, generator=g)
        ul_input = torch.rand(nb, 3, s, s, device=dev, generator=g)
        for o in opts:
            o.zero_grad(set_to_none=True)
        with torch.no_grad():
            for m in models:
                m.eval()
            pseudo_1_score = models[0](ul_input)[0]
            pseudo_2_score = models[1](ul_input)[0]
            for m in models:
                m.train()
        with torch.autocast("cuda", dtype=torch.float16):
            sup1, cl1, _u1 = models[0](l_input)
            sup2, cl2, _u2 = models[1](l_input)
            ul1, cul1, _u3 = models[0](ul_input)
            ul2, cul2, _u4 = models[1](ul_input)
        pred_1, pred_2 = torch.cat([sup1, ul1]), torch.cat([sup2, ul2])
        pseudo_1, pseudo_2 = pred_1.argmax(1), pred_2.argmax(1)
        del pseudo_1_score, pseudo_2_score                     # (they feed the prototype loss in the reference: left out)
        with torch.autocast("cuda", dtype=torch.float16):
            f1, f2 = score_mask(pred_1, pseudo_1, 0.7), score_mask(pred_2, pseudo_2, 0.7)
            cps = 0.5 * ce(pred_1.float(), f2) + 0.5 * ce(pred_2.float(), f1) + dice_loss(pred_1, f2) + dice_loss(pred_2, f1)
            sup = 0.5 * ce(sup1.float(), l_target) + dice_loss(sup1, l_target) + 0.5 * ce(sup2.float(), l_target) + dice_loss(sup2, l_target)
            commitment = (cl1 + cl2 + cul1 + cul2).sum()
            loss = sup + cps + commitment
        scaler.scale(loss).backward()
        for o in opts:
            scaler.step(o)
        scaler.update()
        return loss

    step.scaler = scaler
    return models, opts, step


def run_arm(arm, args, dev, rank, world):
    import torch.distributed as dist
    models, opts, step = build(arm, args, dev, rank, world)
    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    nb = args.per_gpu_batch
    out = {"ms_per_step": ms, "train_img_per_s": nb * world / (ms * 1e-3), "final_loss": float(loss.detach().float().item()),
           "grad_scale_at_end": float(step.scaler.get_scale())}
    del models, opts
    torch.cuda.empty_cache()
    return out


def run_c3(dev, rank, world, steps=10, warmup=6, per_gpu_batch=4, size=512, arms=("b200", "torch", "identity")):
    args = argparse.Namespace(steps=steps, warmup=warmup, per_gpu_batch=per_gpu_batch, size=size)
    res = {"images_per_step": f"{per_gpu_batch * world} labelled + {per_gpu_batch * world} unlabelled at {size}x{size}, "
                              f"two models: 4 train-mode + 2 eval-mode forwards, 1 backward, 2 Adam steps, fp16 autocast",
           "callers": "stand-in resnet50 encoder + U-Net decoder of VQRePTUnet1x1v2's layer shapes (stock PyTorch)",
           "collectives": ("DDP gradient all-reduce (NCCL) + per VQ layer and training forward one all-reduce of counts[K] "
                           "+ sums[K, D] (EMA codebook statistics)") if world > 1 else "none (1 GPU)"}
    for arm in arms:
        res[arm] = run_arm(arm, args, dev, rank, world)
    if "b200" in res and "identity" in res:
        res["vq_ms_per_step_b200"] = res["b200"]["ms_per_step"] - res["identity"]["ms_per_step"]
    if "torch" in res and "identity" in res:
        res["vq_ms_per_step_torch_eager"] = res["torch"]["ms_per_step"] - res["identity"]["ms_per_step"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--per-gpu-batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--arms", default="b200,torch,identity")
    a = ap.parse_args()
    import torch.distributed as dist
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    res = run_c3(dev, rank, world, a.steps, a.warmup, a.per_gpu_batch, a.size, tuple(a.arms.split(",")))
    if rank == 0:
        print(json.dumps({"config": "BASELINE configs[2]: vqreptunet1x1v2 training step", "n_gpus": world, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
