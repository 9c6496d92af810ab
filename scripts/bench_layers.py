"""Secondary measurements on the B200 (informational; the contract line is bench.py):
  * config 1 / 3 layer shapes: VectorQuantizer train-mode forward+backward per layer, ours vs the same
    PyTorch-eager module run ON THE GPU (the reference's own code path on this box: cdist/argmin/one_hot/matmul
    through cuBLAS fp32), vs the CPU numbers of BASELINE.md;
  * config 2 eval forward, same comparison;
  * k-means (config-4-like, reduced N) assign+update iteration.
Run: python scripts/bench_layers.py   (prints a markdown table)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vq_seg_b200 as V  # noqa: E402
from vq_seg_b200 import ops  # noqa: E402
from oracle.vq_oracle import OracleVectorQuantizer  # noqa: E402  (plain torch: also runs on cuda)

dev = torch.device("cuda:0")


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def layer(b, c, h, w, k, train):
    g = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn(b, c, h, w, generator=g)).to(dev)
    e = torch.randn(k, c, generator=g).to(dev)
    ours = V.VectorQuantizer(dim=c, num_embeddings=k).to(dev)
    ref = OracleVectorQuantizer(dim=c, num_embeddings=k).to(dev)
    ref.faithful_ops = True            # the reference's literal op sequence: cdist, argmin, one_hot, matmul, bincount
    for m in (ours, ref):
        m.codebook.embedding.weight.data.copy_(e)
        m.train(train)
    gq = torch.randn(b, c, h, w, device=dev)

    def run(m):
        if train:
            xg = x.clone().requires_grad_(True)
            q, idx, loss, usage = m(xg)
            ((q * gq).sum() + loss.sum()).backward()
            return xg.grad
        with torch.no_grad():
            return m(x)[0]
    t_ours = timeit(lambda: run(ours))
    t_ref = timeit(lambda: run(ref), n=10, warm=2)
    same = torch.equal(ours(x)[1], ref(x)[1])
    return t_ours, t_ref, same


if __name__ == "__main__":
    print("| case | N | D | K | mode | ours (us) | torch-eager on the same B200 (us) | speed-up | same indices |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, (b, c, h, w, k) in {"C2": (8, 256, 64, 64, 512), "C1 level 3": (2, 512, 64, 64, 512),
                                  "C1 level 4": (2, 1024, 32, 32, 512), "C1 level 5": (2, 2048, 16, 16, 512),
                                  "C3 level 3 (B=4)": (4, 512, 64, 64, 512), "C3 level 4 (B=4)": (4, 1024, 32, 32, 512),
                                  "C3 level 5 (B=4)": (4, 2048, 16, 16, 512)}.items():
        for train in (False, True):
            to, tr, same = layer(b, c, h, w, k, train)
            print(f"| {name} | {b*h*w} | {c} | {k} | {'train fwd+bwd' if train else 'eval fwd'} | {to:.0f} | {tr:.0f} | {tr/to:.1f}x | {same} |", flush=True)
    # k-means iteration, config-4-like at reduced N (row-major samples)
    for (n, d, k) in [(200_000, 512, 1024), (1_000_000, 512, 1024)]:
        x = torch.randn(1, n, d, device=dev)
        init = torch.arange(k, device=dev)
        t = timeit(lambda: V.kmeans(x, k, 1, init_indices=init), n=5, warm=1)
        print(f"| k-means 1 iteration | {n} | {d} | {k} | assign + ordered stats + finalize | {t:.0f} | - | - | - |", flush=True)
        t = timeit(lambda: V.kmeans(x, k, 1, init_indices=init, deterministic=False), n=5, warm=1)
        print(f"| k-means 1 iteration (atomic stats) | {n} | {d} | {k} | assign + atomic stats + finalize | {t:.0f} | - | - | - |", flush=True)
