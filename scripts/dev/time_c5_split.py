"""Developer tool: filter / rescoring split of the config-5 assignment (4 Mi x 256 against K = 65536) and of a config-4
slice, with the share of rows the filter leaves undecided."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops, _native
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
def run(name, n, d, k):
    x = torch.randn(1, n, d, generator=g, device=dev)
    e = x[0, torch.randperm(n, generator=g, device=dev)[:k]].contiguous() + 0.1 * torch.randn(k, d, generator=g, device=dev)
    blob = ops.prepare_codebook(e)
    prof = _native.ProfileEvents(); ops.set_profile_events(prof)
    for _ in range(3):
        idx, counts = ops.assign(x, e, blob, ops.ALGO_AUTO); torch.cuda.synchronize()
    ops.set_profile_events(None)
    ws = ops._last_assign_ws
    print(f"{name}: undecided {ws[:4].view(torch.int32).item()} of {n} rows, overflow {ws[4:8].view(torch.int32).item()}; "
          f"filter {prof.filter_ms():.2f} ms, rescoring + overflow {prof.rescore_ms():.2f} ms", flush=True)
run("config 5 (4 Mi x 256, K = 65536)", 1 << 22, 256, 65536)
run("config 4 slice (2 Mi x 512, K = 1024)", 1 << 21, 512, 1024)
run("K = 8192, D = 256, 2 Mi rows", 1 << 21, 256, 8192)
