"""Developer tool: cost of a handful of overflow rows at K = 65536 (the rescoring + overflow share of one assignment)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops, _native
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
n, d, k = 1 << 18, 256, 65536
x = torch.randn(1, n, d, generator=g, device=dev)
e = torch.randn(k, d, generator=g, device=dev)
e[100:112] = 0.8 * float(e.norm(dim=-1).min()) * torch.nn.functional.normalize(e[100:112], dim=-1)
for zero_rows in (0, 1, 3):
    xx = x.clone()
    if zero_rows: xx[0, :zero_rows] = 0.0
    blob = ops.prepare_codebook(e)
    prof = _native.ProfileEvents(); ops.set_profile_events(prof)
    for _ in range(3):
        idx, counts = ops.assign(xx, e, blob, ops.ALGO_AUTO); torch.cuda.synchronize()
    ops.set_profile_events(None)
    ws = ops._last_assign_ws
    print(f"{zero_rows} zero rows: overflow rows {ws[4:8].view(torch.int32).item()}, filter {prof.filter_ms():.2f} ms, rescoring + overflow {prof.rescore_ms():.3f} ms", flush=True)
