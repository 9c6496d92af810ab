"""Developer tool: the lookup on degenerate codebooks (the reference's default uniform(-1/K, 1/K) init against
real-sized features; a collapsed codebook of near-duplicates), where the fp16 filter cannot separate the codes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops, _native
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
def run(name, xv, e):
    blob = ops.prepare_codebook(e)
    prof = _native.ProfileEvents(); ops.set_profile_events(prof)
    for _ in range(3):
        idx, counts = ops.assign(xv, e, blob, ops.ALGO_TC); torch.cuda.synchronize()
    ops.set_profile_events(None)
    ws = ops._last_assign_ws
    n = xv.shape[0] * xv.shape[1]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d = torch.cdist(xv.contiguous(), e.unsqueeze(0).expand(xv.shape[0], -1, -1)); ref = d.argmin(-1); b.record(); torch.cuda.synchronize()
    print(f"{name}: undecided {ws[:4].view(torch.int32).item()} of {n}, overflow rows {ws[4:8].view(torch.int32).item()}; filter {prof.filter_ms()*1e3:.0f} us, "
          f"rescoring + overflow {prof.rescore_ms()*1e3:.0f} us; torch cdist+argmin on the same GPU {a.elapsed_time(b)*1e3:.0f} us; "
          f"indices equal torch's on {100.0 * (ref == idx).float().mean().item():.2f} % of rows")
x = torch.relu(torch.randn(8, 256, 4096, generator=g, device=dev)); xv = x.permute(0, 2, 1)
e = (torch.rand(512, 256, generator=g, device=dev) * 2 - 1) / 512
run("C2 map, uniform(-1/K, 1/K) codebook (the reference's default init)", xv, e)
e2 = xv.reshape(-1, 256)[:8].repeat(64, 1).contiguous() + 1e-7 * torch.randn(512, 256, generator=g, device=dev)
run("C2 map, collapsed codebook (8 distinct codes x 64 near-copies)", xv, e2)
e3 = xv.reshape(-1, 256)[torch.randperm(32768, device=dev)[:512]].contiguous()
run("C2 map, codes = rows of the map", xv, e3)
