"""Developer diagnostic: how do the tcgen05 filter and the rescoring pass behave on the feature maps of the config-3
stand-in model (untrained resnet50 features, k-means-initialised codebooks)?  Prints per VQ layer the share of rows the
filter could not decide, the share whose short-list overflowed, and the kernel times."""
import os, sys, argparse
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import train_step_c3 as T
import vq_seg_b200 as V
from vq_seg_b200 import ops, _native

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 0
arm = sys.argv[2] if len(sys.argv) > 2 else "b200_noema"
args = argparse.Namespace(steps=1, warmup=1, per_gpu_batch=4, size=512)
models, opts, step = T.build(arm, args, dev, 0, 1)
model = models[0]
model.train()
g = torch.Generator(device=dev).manual_seed(100)
x = torch.rand(4, 3, 512, 512, device=dev, generator=g)
if n_steps == 0:
    with torch.autocast("cuda", dtype=torch.float16):
        out = model(x)                      # k-means init happens here
for i in range(n_steps):
    loss = step()
print(f"after {n_steps} training steps of arm {arm}")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    feats = model.encoder(x)[1:]
for i, m in enumerate(model.codebook):
    if not isinstance(m, V.VectorQuantizer):
        continue
    f = feats[i].float()
    b, c, h, w = f.shape
    xv = f.reshape(b, c, h * w).permute(0, 2, 1)
    e = m.codebook.embedding.weight.detach()
    blob = ops.prepare_codebook(e)
    prof = _native.ProfileEvents()
    ops.set_profile_events(prof)
    for _ in range(3):
        idx, counts = ops.assign(xv, e, blob, ops.ALGO_TC)
        torch.cuda.synchronize()
    ops.set_profile_events(None)
    ws = ops._last_assign_ws
    n = b * h * w
    flagged = ws[:4].view(torch.int32).item()
    recs = ws[256 + 8192:256 + 8192 + 48 * n].view(torch.int32).reshape(n, 12)[:flagged]
    cnt = recs[:, 1]
    over = int((cnt > 8).sum())
    i_ex, _ = ops.assign(xv, e, None, ops.ALGO_EXACT)
    d = torch.cdist(xv.reshape(1, -1, c), e.unsqueeze(0))[0]
    top2 = d.topk(2, dim=-1, largest=False).values
    print(f"layer {i}: N={n} D={c} K={e.shape[0]}  |x| mean {xv.norm(dim=-1).mean().item():.3g}  |e| mean {e.norm(dim=-1).mean().item():.3g}  "
          f"d1 mean {top2[:, 0].mean().item():.3g}  (d2-d1)/d1 median {((top2[:, 1] - top2[:, 0]) / top2[:, 0].clamp_min(1e-20)).median().item():.3g}")
    print(f"    undecided {100.0 * flagged / n:.1f}% of rows, short-list overflow {100.0 * over / n:.1f}%, mean candidates of the listed "
          f"{cnt[cnt <= 8].float().mean().item() if (cnt <= 8).any() else 0:.2f}; filter {prof.filter_ms() * 1e3:.1f} us, rescoring {prof.rescore_ms() * 1e3:.1f} us; "
          f"tc == exact: {bool(torch.equal(idx, i_ex))}; codes used {(counts > 0).sum().item()}")
    if c == 1024:
        def timed(xx, ee, bb, label):
            prof = _native.ProfileEvents(); ops.set_profile_events(prof)
            for _ in range(3):
                ops.assign(xx, ee, bb, ops.ALGO_TC); torch.cuda.synchronize()
            ops.set_profile_events(None)
            fl = ops._last_assign_ws[:4].view(torch.int32).item()
            print(f"    [{label}] undecided {fl}, filter {prof.filter_ms() * 1e3:.1f} us, rescoring {prof.rescore_ms() * 1e3:.1f} us")
        print("    cnt histogram of the undecided rows:", torch.bincount(cnt.long(), minlength=10).tolist())
        rows_of = recs[:, 0].long()
        print("    undecided rows per image:", torch.bincount(rows_of // (h * w), minlength=b).tolist(), " distinct first candidates:", recs[:, 4].unique().numel())
        timed(xv, e, blob, "as is (NCHW)")
        xc = xv.contiguous()
        timed(xc, e, blob, "packed rows")
        perm = torch.randperm(e.shape[0], device=dev)
        e2 = e[perm].contiguous()
        timed(xv, e2, ops.prepare_codebook(e2), "codes shuffled")
        xr = torch.relu(torch.randn(b, c, h * w, device=dev)).permute(0, 2, 1) * 3
        e3 = xr.reshape(-1, c)[torch.randperm(n, device=dev)[:512]].contiguous()
        timed(xr, e3, ops.prepare_codebook(e3), "synthetic relu-randn, codes = rows")
