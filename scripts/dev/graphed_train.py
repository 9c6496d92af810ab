"""Developer tool: VectorQuantizer training forward + backward as torch.cuda.make_graphed_callables against eager."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vq_seg_b200 as V
dev = torch.device("cuda:0")
for shape in ((2, 512, 64, 64), (2, 1024, 32, 32), (2, 2048, 16, 16), (8, 256, 64, 64)):
    torch.manual_seed(0)
    m = V.VectorQuantizer(dim=shape[1], num_embeddings=512).to(dev).train()
    x = torch.randn(*shape, device=dev).relu_().requires_grad_(True)
    gy = torch.randn(*shape, device=dev)
    def step(fn, xin):
        q, idx, loss, usage = fn(xin)
        (q * gy).sum().add(loss.sum()).backward()
        return q, idx, loss, usage
    q0, i0, l0, u0 = step(m, x); g0 = x.grad.clone(); x.grad = None
    xs = x.detach().clone().requires_grad_(True)
    gm = torch.cuda.make_graphed_callables(m, (xs,), allow_unused_input=True)
    q1, i1, l1, u1 = step(gm, x); g1 = x.grad.clone(); x.grad = None
    print(shape, "equal:", bool(torch.equal(q0, q1)), bool(torch.equal(i0, i1)), bool(torch.equal(l0, l1)), bool(torch.equal(u0, u1)), bool(torch.equal(g0, g1)))
    def t(fn, n=200):
        for _ in range(20): step(fn, x); x.grad = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): step(fn, x); x.grad = None
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    print("   eager %.1f us, graphed %.1f us per train forward + backward" % (t(m), t(gm)), flush=True)
