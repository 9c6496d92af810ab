"""Developer tool: the tensor-core filters against the brute-force exact scorer over a sweep of operand scales (the
filter's slack must stay a rigorous bound when |e| << |x|, |e| >> |x|, relu / signed inputs, several widths)."""
import os, sys, itertools, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vq_seg_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(3)
bad = 0
for (D, K, P), xs, es, relu, nchw in itertools.product(((256, 512, 16384), (64, 512, 16384), (512, 1024, 8192), (1024, 512, 4096), (128, 4096, 8192)),
                                                        (1e-3, 1.0, 40.0), (1e-4, 1.0 / 512, 1.0, 40.0), (False, True), (True, False)):
    x = torch.randn(4, D, P, generator=g, device=dev) * xs
    if relu: x = torch.relu(x)
    xv = x.permute(0, 2, 1) if nchw else x.permute(0, 2, 1).contiguous()
    e = ((torch.rand(K, D, generator=g, device=dev) * 2 - 1) * es).contiguous()
    blob = ops.prepare_codebook(e)
    i_tc, c_tc = ops.assign(xv, e, blob, ops.ALGO_AUTO)
    i_ex, c_ex = ops.assign(xv, e, None, ops.ALGO_EXACT)
    ws = ops._last_assign_ws
    ok = bool(torch.equal(i_tc, i_ex) and torch.equal(c_tc, c_ex))
    if not ok:
        bad += 1
        print("MISMATCH", D, K, P, xs, es, relu, nchw, int((i_tc != i_ex).sum()), flush=True)
print("cases with a mismatch:", bad)
