"""Developer diagnostics on the GPU box (not a test, not the bench): python scripts/gpu_dev.py <stage>.
Each stage prints PASS/FAIL lines and timings; run under `timeout` from gpurun."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from oracle import vq_oracle as O  # noqa: E402
import vq_seg_b200 as V  # noqa: E402
from vq_seg_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "golden_v1.pt"), weights_only=False)


def view(x):
    b, c, h, w = x.shape
    return x.reshape(b, c, h * w).permute(0, 2, 1)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3      # us


def stage_exact():
    names = sys.argv[2:] or list(cases.FORWARD_CASES)
    for name in names:
        x, e = cases.FORWARD_CASES[name]()
        rec = GOLD["forward"][name]
        xd, ed = x.to(dev), e.to(dev)
        idx, counts = ops.assign(view(xd), ed, None, ops.ALGO_EXACT)
        torch.cuda.synchronize()
        gi = rec["idx"].reshape(idx.shape).to(torch.int64)
        mism = (idx.cpu() != gi).sum().item()
        cm = (counts.cpu() != rec["counts"]).sum().item()
        print(f"[exact] {name:16s} idx mismatches {mism}/{gi.numel()} counts mism {cm}  {'PASS' if mism == 0 and cm == 0 else 'FAIL'}", flush=True)


def stage_tc():
    names = sys.argv[2:] or list(cases.FORWARD_CASES)
    for name in names:
        x, e = cases.FORWARD_CASES[name]()
        rec = GOLD["forward"][name]
        xd, ed = x.to(dev), e.to(dev)
        blob = ops.prepare_codebook(ed)
        torch.cuda.synchronize()
        idx_e, counts_e = ops.assign(view(xd), ed, None, ops.ALGO_EXACT)
        idx, counts = ops.assign(view(xd), ed, blob, ops.ALGO_TC)
        torch.cuda.synchronize()
        gi = rec["idx"].reshape(idx.shape).to(torch.int64)
        mism = (idx.cpu() != gi).sum().item()
        mism_e = (idx != idx_e).sum().item()
        cm = (counts != counts_e).sum().item()
        ws = ops._last_assign_ws
        flagged = ws[:4].view(torch.int32).item()
        nrows = gi.numel()
        # workspace: 256 B counters | 8 KiB overflow-split scratch | one 48-byte record per undecided row {row, n, -, -, cand[8]}
        recs = ws[256 + 8192:256 + 8192 + 48 * nrows].view(torch.int32).reshape(nrows, 12)[:flagged]
        hist = torch.bincount(recs[:, 1].long().clamp(0, 9), minlength=10).tolist() if flagged else []
        print("      candidate-count histogram of rescored rows (9 = all codes):", hist)
        print(f"[tc] {name:16s} vs golden {mism}/{gi.numel()}  vs exact {mism_e}  counts mism {cm}  rescored rows {flagged} ({100.0 * flagged / gi.numel():.1f}%)  {'PASS' if mism_e == 0 and cm == 0 else 'FAIL'}", flush=True)


def stage_ops():
    for name in ["d64", "c1_l4", "r448_l5", "odd_7x7", "c2_randn"]:
        x, e = cases.FORWARD_CASES[name]()
        rec = GOLD["forward"][name]
        m = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0]).to(dev)
        m.codebook.embedding.weight.data.copy_(e)
        m.codebook.algo = ops.ALGO_EXACT
        m.eval()
        with torch.no_grad():
            q, idx, loss, usage = m(x.to(dev))
        ok_e = cases.sha(q.cpu()) == rec["q_eval_sha"] and torch.equal(usage.cpu(), rec["usage"]) and torch.equal(loss.cpu(), rec["loss_eval"])
        m.train()
        xg = x.to(dev).requires_grad_(True)
        q, idx, loss, usage = m(xg)
        ok_t = cases.sha(q.detach().cpu()) == rec["q_train_sha"]
        rel = ((loss.detach().cpu() - rec["loss_train"]).abs() / rec["loss_train"].abs()).item()
        g = torch.Generator().manual_seed(999)
        gq = torch.randn(q.shape, generator=g).to(dev)
        ((q * gq).sum() + 1.5 * loss.sum()).backward()
        ref = O.vq_backward(x, O.vq_forward(x, e, True, 1)["quantize"], gq.cpu(), torch.tensor(1.5), 1)
        gerr = ((xg.grad.cpu() - ref).abs().max() / ref.abs().max()).item()
        print(f"[ops] {name:12s} eval-bitexact {ok_e} train-q-bitexact {ok_t} loss rel {rel:.2e} grad relmax {gerr:.2e} "
              f"wgrad None {m.codebook.embedding.weight.grad is None} loss.shape {tuple(loss.shape)} rg {loss.requires_grad}", flush=True)
    for name in cases.KMEANS_CASES:
        x, k, iters, init_idx, use_cos = cases.kmeans_case(name)
        rec = GOLD["kmeans"][name]
        xv = view(x.to(dev))
        src = ops.l2norm_rows(xv) if use_cos else xv
        means, bins = V.kmeans(src, k, iters, use_cosine_sim=use_cos, init_indices=init_idx, algo=ops.ALGO_EXACT)
        torch.cuda.synchronize()
        be = torch.equal(bins[0].cpu(), rec["bins"])
        me = torch.equal(means[0].cpu(), rec["means"])
        err = (means[0].cpu() - rec["means"]).abs().max().item()
        print(f"[kmeans] {name:12s} bins exact {be} means bit-exact {me} maxabs {err:.3e}", flush=True)


def stage_time():
    x, e = cases.FORWARD_CASES["c2_randn"]()
    xd, ed = x.to(dev), e.to(dev)
    xv = view(xd)
    blob = ops.prepare_codebook(ed)
    from vq_seg_b200 import _native
    L = _native.lib()
    for algo, nm in [(ops.ALGO_EXACT, "exact"), (ops.ALGO_TC, "tc+rescore")]:
        if algo == ops.ALGO_EXACT and "noexact" in sys.argv:
            continue
        med, best = timeit(lambda: ops.assign(xv, ed, blob, algo), n=10 if algo == ops.ALGO_EXACT else 50)
        print(f"[time] assign {nm:10s} C2 (x L2-resident): median {med:9.1f} us best {best:9.1f} us -> {2*32768*512*256/med/1e6:.1f} TFLOP/s", flush=True)
    L.vqseg_set_kernel_timing(1)
    kt, rt = [], []
    for _ in range(20):
        ops.assign(xv, ed, blob, ops.ALGO_TC); torch.cuda.synchronize()
        kt.append(L.vqseg_get_kernel_timing_ms(0) * 1e3); rt.append(L.vqseg_get_kernel_timing_ms(1) * 1e3)
    # cold-L2 variant: thrash L2 between calls
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kc, rc = [], []
    for _ in range(20):
        junk.fill_(1); ops.assign(xv, ed, blob, ops.ALGO_TC); torch.cuda.synchronize()
        kc.append(L.vqseg_get_kernel_timing_ms(0) * 1e3); rc.append(L.vqseg_get_kernel_timing_ms(1) * 1e3)
    L.vqseg_set_kernel_timing(0)
    kt.sort(); rt.sort(); kc.sort(); rc.sort()
    print(f"[time] filter kernel  : warm-L2 median {kt[10]:.1f} us (best {kt[0]:.1f}); cold-L2 median {kc[10]:.1f} us (best {kc[0]:.1f})")
    print(f"[time] rescore kernel : warm-L2 median {rt[10]:.1f} us (best {rt[0]:.1f}); cold-L2 median {rc[10]:.1f} us (best {rc[0]:.1f})", flush=True)
    idx, _ = ops.assign(xv, ed, blob, ops.ALGO_EXACT)
    med, best = timeit(lambda: ops.prepare_codebook(ed), n=50)
    print(f"[time] prepare_codebook: median {med:.1f} us best {best:.1f}")
    med, best = timeit(lambda: ops.gather_ste(xv, ed, idx, ops.MODE_TRAIN), n=50)
    print(f"[time] gather_ste train: median {med:.1f} us best {best:.1f} -> {(8*32768*256+8*32768)/med/1e3:.1f} GB/s")
    q, _ = ops.gather_ste(xv, ed, idx, ops.MODE_TRAIN)
    gq = torch.randn_like(q)
    gm = torch.ones(1, device=dev)
    med, best = timeit(lambda: ops.ste_bwd(gq, xv, q, gm, 2.0 / xv.numel()), n=50)
    print(f"[time] ste_bwd: median {med:.1f} us best {best:.1f} -> {(16*32768*256)/med/1e3:.1f} GB/s")
    for det in (True, False):
        med, best = timeit(lambda: ops.code_stats(xv, idx, 512, det), n=20)
        print(f"[time] code_stats det={det}: median {med:.1f} us best {best:.1f}")


def stage_prof():
    """one launch of every kernel at C2 (for `ncu --metrics gpu__time_duration.sum`)"""
    x, e = cases.FORWARD_CASES["c2_randn"]()
    xd, ed = x.to(dev), e.to(dev)
    xv = view(xd)
    for rep in range(3):
        blob = ops.prepare_codebook(ed)
        idx, counts = ops.assign(xv, ed, blob, ops.ALGO_TC)
        q, mse = ops.gather_ste(xv, ed, idx, ops.MODE_TRAIN)
        gx = ops.ste_bwd(torch.ones_like(q), xv, q, torch.ones(1, device=dev), 2.0 / xv.numel())
        c, s = ops.code_stats(xv, idx, 512, True)
        c2, s2 = ops.code_stats(xv, idx, 512, False)
        u = ops.code_usage(counts)
        torch.cuda.synchronize()
    print("prof done", idx.sum().item(), mse.item(), u.item())


def stage_bw():
    from vq_seg_b200 import _native
    L = _native.lib()
    x = torch.randn(8 * 256 * 4096, device=dev)          # 33.5 MB, rows of 4096 floats (one image's d-row)
    sink = torch.zeros(4, device=dev)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (pat, depth, nm) in [(0, 8, "LDG.128 8x64B/instr"), (1, 4, "LDG.256 8x128B/instr"), (2, 8, "LDG.128 512B contiguous"),
                             (22, 8, "contiguous, 2 CTA/SM"), (42, 8, "contiguous, 4 CTA/SM"), (82, 8, "contiguous, 8 CTA/SM")]:
        for cold in (False, True):
            ts = []
            for _ in range(10):
                if cold:
                    junk.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = L.vqseg_debug_load_bandwidth(x.data_ptr(), x.numel(), 4096, pat, depth, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
                b.record(); torch.cuda.synchronize()
                assert rc == 0
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            print(f"[bw] {nm:26s} {'cold' if cold else 'L2-warm':8s}: median {ts[5]:7.1f} us -> {x.numel() * 4 / ts[5] / 1e6:6.2f} TB/s (best {ts[0]:.1f} us)")
    big = torch.randn(1 << 28, device=dev)     # 1 GiB
    for (pat, depth, nm) in [(0, 8, "P0 LDG.128 8x64B depth 8"), (0, 24, "P0 LDG.128 8x64B depth 24"), (1, 4, "P1 LDG.256 8x128B depth 4"),
                             (1, 12, "P1 LDG.256 8x128B depth 12"), (2, 8, "P2 contiguous depth 8"), (2, 24, "P2 contiguous depth 24"),
                             (82, 8, "P2 contiguous 8 CTA/SM")]:
        for buf, bn in ((big, "1 GiB"), (x, "33.5 MB")):
            ts = []
            for _ in range(5):
                if bn != "1 GiB":
                    junk.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = L.vqseg_debug_load_bandwidth(buf.data_ptr(), buf.numel(), 4096, pat, depth, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
                b.record(); torch.cuda.synchronize()
                assert rc == 0, rc
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            print(f"[bw2] {nm:30s} {bn:8s}: median {ts[2]:8.1f} us -> {buf.numel() * 4 / ts[2] / 1e6:6.2f} TB/s")


def stage_bwbulk():
    """bulk-copy (cp.async.bulk) loader vs register loaders: can one CTA per SM stream x at HBM rate?"""
    from vq_seg_b200 import _native
    L = _native.lib()
    x = torch.randn(8 * 256 * 4096, device=dev)
    big = torch.randn(1 << 28, device=dev)
    sink = torch.zeros(4, device=dev)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (pat, depth, nm) in [(0, 24, "registers LDG.128 depth 24"), (2, 24, "registers contiguous d24"), (82, 8, "registers contig 8 CTA/SM"),
                             (3, 4, "bulk 512B strided, 4 st"), (3, 16, "bulk 512B strided, 16 st"),
                             (4, 4, "bulk 8KiB linear, 4 st"), (4, 16, "bulk 8KiB linear, 16 st"),
                             (5, 4, "bulk 512B, 4 warps, 4 st"), (5, 16, "bulk 512B, 4 warps, 16 st")]:
        for buf, bn in ((big, "1 GiB"), (x, "33.5 MB")):
            ts = []
            for _ in range(7):
                if bn != "1 GiB":
                    junk.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = L.vqseg_debug_load_bandwidth(buf.data_ptr(), buf.numel(), 4096, pat, depth, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
                b.record(); torch.cuda.synchronize()
                assert rc == 0, rc
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            print(f"[bwbulk] {nm:28s} {bn:8s}: median {ts[3]:8.1f} us -> {buf.numel() * 4 / ts[3] / 1e6:6.2f} TB/s (best {ts[0]:.1f})")


def stage_seghead():
    """VQ segmentation head at the reference's decoder-output size (resnet50 U-Net, 4 x 32 x 256 x 256, 3 classes):
    map kernel and its backward by CUDA events, module forward / forward+backward against the same module written
    with torch ops (the oracle restatement moved to the GPU)."""
    import vq_seg_b200 as V
    from oracle.seghead_oracle import OracleVQSegmentationHead
    b, c, h, w, k = 4, 32, 256, 256, 3
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.relu(torch.randn(b, c, h, w, generator=g, device=dev)) for _ in range(6)]     # 6 x 33.5 MB > L2
    e = torch.rand(k, c, generator=g, device=dev)

    def timeit(fn, n=30):
        for i in range(3):
            fn(xs[i % len(xs)])
        torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(xs[i % len(xs)])
        bb.record(); torch.cuda.synchronize()
        return a.elapsed_time(bb) * 1e3 / n

    def timeit_graph(fn, n=12):
        """n calls (cycling the inputs) captured into one CUDA graph: no host time between the launches"""
        for i in range(3):
            fn(xs[i % len(xs)])
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(n):
                fn(xs[i % len(xs)])
        gr.replay(); torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gr.replay()
        bb.record(); torch.cuda.synchronize()
        return a.elapsed_time(bb) * 1e3 / (5 * n)

    view_ = lambda x: x.reshape(b, c, h * w).permute(0, 2, 1)
    t_map = timeit_graph(lambda x: ops._dist_map_impl(view_(x), e, False))
    dist, idx, counts = ops._dist_map_impl(view_(xs[0]), e, False)
    gd = torch.randn_like(dist)
    t_bwd = timeit_graph(lambda x: ops._dist_map_bwd_impl(gd, dist, view_(x), e))
    n = b * h * w
    by_f = 4 * n * c + 4 * n * k + 8 * n
    by_b = 4 * n * c * 2 + 8 * n * k
    print(f"[seghead] map kernel   : {t_map:7.1f} us  ({by_f / t_map / 1e6:5.2f} TB/s of {by_f / 1e6:.1f} MB algorithmic)")
    print(f"[seghead] map backward : {t_bwd:7.1f} us  ({by_b / t_bwd / 1e6:5.2f} TB/s of {by_b / 1e6:.1f} MB algorithmic)")
    ours = V.VQSegmentationHead(dim=c, num_embeddings=k).to(dev)
    ref = OracleVQSegmentationHead(dim=c, num_embeddings=k).to(dev)
    ours.codebook.embedding.weight.data.copy_(e); ref.embedding.weight.data.copy_(e)
    for nm, m in (("vq_seg_b200", ours), ("torch ops (oracle module on the GPU)", ref)):
        m.eval()
        with torch.no_grad():
            t_eval = timeit(lambda x: m(x))
        m.train()

        def step(x):
            xg = x.detach().requires_grad_(True)
            q, score, i_, loss, u = m(xg)
            (score.sum() + q.sum() + loss.sum()).backward()
        t_train = timeit(step, n=10)
        print(f"[seghead] {nm:38s}: eval forward {t_eval:8.1f} us, train forward+backward {t_train:8.1f} us")


def stage_big():
    """BASELINE configs 4 and 5 on one GPU: k-means iteration at N = 1e7, K = 1024, D = 512 (20 GB of samples) and a
    full-codebook assignment at N = 4 Mi, K = 65536, D = 256 (140.7 TFLOP)."""
    def ev_time(fn, n=1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            r = fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n, r
    g = torch.Generator(device="cuda").manual_seed(0)
    n, d, k = 10_000_000, 512, 1024
    x = torch.empty(1, n, d, device=dev)
    for i in range(0, n, 1_000_000):
        x[0, i:i + 1_000_000].normal_(generator=g)
    means = x[0, :k].clone()
    blob = ops.prepare_codebook(means)
    ops.assign(x, means, blob, ops.ALGO_TC)                                   # warm-up
    t_as, (idx, counts) = ev_time(lambda: ops.assign(x, means, blob, ops.ALGO_TC))
    t_at, _ = ev_time(lambda: ops.code_stats(x, idx, k, False))
    t_dt, (bins, sums) = ev_time(lambda: ops.code_stats(x, idx, k, True))
    assert counts.sum().item() == n and torch.equal(bins, counts)
    fl = 2.0 * n * k * d
    print(f"[big] C4 N=1e7 K=1024 D=512: assign {t_as:.2f} ms ({fl / t_as / 1e9:.0f} TFLOP/s), stats atomic {t_at:.2f} ms "
          f"({(4.0 * n * d) / t_at / 1e9:.2f} TB/s), stats ordered {t_dt:.2f} ms -> iteration {t_as + t_at:.1f} / {t_as + t_dt:.1f} ms")
    del x, idx, sums
    torch.cuda.empty_cache()
    n, d, k = 4 << 20, 256, 65536
    x = torch.randn(1, n, d, generator=g, device=dev)
    e = torch.randn(k, d, generator=g, device=dev)
    blob = ops.prepare_codebook(e)
    t5, (idx, counts) = ev_time(lambda: ops.assign(x, e, blob, ops.ALGO_TC))
    fl = 2.0 * n * k * d
    print(f"[big] C5 N=4Mi K=65536 D=256 (whole codebook on one GPU): assign {t5:.1f} ms ({fl / t5 / 1e9:.0f} TFLOP/s), "
          f"codes used {int((counts > 0).sum())}")
    # spot check of 2048 rows against the exact scorer
    sub = x[:, 12345:12345 + 2048]
    i_ex, _ = ops.assign(sub, e, None, ops.ALGO_EXACT)
    assert torch.equal(i_ex, idx[:, 12345:12345 + 2048]), "C5 spot check failed"
    print("[big] C5 spot check vs exact scorer: ok")


def stage_shapes():
    """filter / rescoring kernel times (CUDA events inside the C ABI) over the BASELINE shapes"""
    from vq_seg_b200 import _native
    shapes = {"C2": (8, 256, 64, 64, 512), "C1 l3": (2, 512, 64, 64, 512), "C1 l4": (2, 1024, 32, 32, 512),
              "C1 l5": (2, 2048, 16, 16, 512), "C3 l3": (4, 512, 64, 64, 512), "C3 l4": (4, 1024, 32, 32, 512),
              "C3 l5": (4, 2048, 16, 16, 512), "r448 l3": (4, 512, 56, 56, 512), "K4096": (8, 256, 64, 64, 4096),
              "C5-like K=16384": (4, 256, 64, 64, 16384), "C4-like rows": (1, 512, 1, 262144, 1024)}
    for nm, (b, c, h, w, k) in shapes.items():
        g = torch.Generator(device="cuda").manual_seed(3)
        if nm.startswith("C4"):
            x = torch.randn(1, h * w, c, generator=g, device=dev)           # packed (N, D) rows
            xv = x
        else:
            x = torch.relu(torch.randn(b, c, h * w, generator=g, device=dev))
            xv = x.permute(0, 2, 1)
        e = torch.randn(k, c, generator=g, device=dev)
        blob = ops.prepare_codebook(e)
        for _ in range(3):
            ops.assign(xv, e, blob, ops.ALGO_TC)
        prof = _native.ProfileEvents()
        ops.set_profile_events(prof)
        kt, rt = [], []
        for _ in range(10):
            ops.assign(xv, e, blob, ops.ALGO_TC); torch.cuda.synchronize()
            kt.append(prof.filter_ms() * 1e3); rt.append(prof.rescore_ms() * 1e3)
        ops.set_profile_events(None)
        kt.sort(); rt.sort()
        n = xv.shape[0] * xv.shape[1]
        flagged = ops._last_assign_ws[:4].view(torch.int32).item()
        fl = 2.0 * n * k * c
        print(f"[shapes] {nm:16s} N={n:7d} D={c:5d} K={k:6d}: filter {kt[5]:8.1f} us ({fl / kt[5] / 1e6:7.1f} TFLOP/s)  rescoring {rt[5]:7.1f} us  rescored {100.0 * flagged / n:5.1f}%", flush=True)


def stage_stream():
    """the two streaming filters (single CTA / TMA-fed CTA pair) side by side on the k-means and large-codebook shapes"""
    from vq_seg_b200 import _native
    shapes = {"C4 rows 256k": ("rows", 262144, 512, 1024), "C4 rows 1M": ("rows", 1 << 20, 512, 1024),
              "C5 rows 64k": ("rows", 65536, 256, 65536), "K4096 nchw": ("nchw", 32768, 256, 4096),
              "K16384 nchw": ("nchw", 16384, 256, 16384), "D128 K2304 rows": ("rows", 262144, 128, 2304),
              "D512 K512 nchw": ("nchw", 65536, 512, 512)}
    sel = sys.argv[2:]
    for nm, (lay, n, d, k) in shapes.items():
        if sel and not any(s in nm for s in sel):
            continue
        g = torch.Generator(device="cuda").manual_seed(3)
        if lay == "rows":
            xv = torch.randn(1, n, d, generator=g, device=dev)
        else:
            xv = torch.relu(torch.randn(4, d, n // 4, generator=g, device=dev)).permute(0, 2, 1)
        e = torch.randn(k, d, generator=g, device=dev)
        blob = ops.prepare_codebook(e)
        res = {}
        samples = ops.prepare_samples(xv) if d <= 512 else None
        for an, algo in (("single", ops.ALGO_TC_STREAM), ("pair", ops.ALGO_TC_STREAM_PAIR), ("prepared", ops.ALGO_AUTO)):
            smp = samples if an == "prepared" else None
            if an == "prepared" and samples is None:
                continue
            for _ in range(2):
                out = ops.assign(xv, e, blob, algo, 0, smp)
            prof = _native.ProfileEvents()
            ops.set_profile_events(prof)
            kt, rt = [], []
            for _ in range(5):
                out = ops.assign(xv, e, blob, algo, 0, smp); torch.cuda.synchronize()
                kt.append(prof.filter_ms() * 1e3); rt.append(prof.rescore_ms() * 1e3)
            ops.set_profile_events(None)
            kt.sort(); rt.sort()
            flagged = ops._last_assign_ws[:4].view(torch.int32).item()
            res[an] = (kt[2], rt[2], flagged, out)
        fl = 2.0 * n * k * d
        same = torch.equal(res["single"][3][0], res["pair"][3][0]) and ("prepared" not in res or torch.equal(res["pair"][3][0], res["prepared"][3][0]))
        prep = f"prepared {res['prepared'][0]:9.1f} us ({fl / res['prepared'][0] / 1e6:6.1f} TF)  " if "prepared" in res else ""
        print(f"[stream] {nm:16s} N={n:8d} D={d:4d} K={k:6d}: single {res['single'][0]:9.1f} us ({fl / res['single'][0] / 1e6:6.1f} TF)  "
              f"pair {res['pair'][0]:9.1f} us ({fl / res['pair'][0] / 1e6:6.1f} TF)  {prep}rescoring {res['pair'][1]:7.1f} us  "
              f"rescored {100.0 * res['pair'][2] / n:4.1f}%  same idx {same}", flush=True)


def stage_null():
    from vq_seg_b200 import _native
    L = _native.lib()
    st = torch.cuda.current_stream().cuda_stream
    small = torch.zeros(1024, device=dev)
    for kind, nm in [(0, "148x640 threads"), (1, "+ 229 KB dynamic smem"), (2, "+ cluster of 2"), (3, "+ TMEM alloc/dealloc + cluster sync")]:
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(200):
                rc = L.vqseg_debug_null_launch(kind, st)
            b.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            ts.append(a.elapsed_time(b) * 1e3 / 200)
        ts.sort()
        print(f"[null] {nm:38s}: {ts[2]:6.2f} us per launch (200 back-to-back launches, GPU queue full)")


def stage_stats():
    """one k-means iteration's kernels at a config-4-like size (packed rows), for ncu's launch list"""
    n, d, k = 1_000_000, 512, 1024
    x = torch.randn(1, n, d, device=dev)
    means = x[0, :k].contiguous()
    blob = ops.prepare_codebook(means)
    for rep in range(2):
        idx, counts = ops.assign(x, means, blob, ops.ALGO_AUTO)
        c1, s1 = ops.code_stats(x, idx, k, True)
        c2, s2 = ops.code_stats(x, idx, k, False)
        torch.cuda.synchronize()
    print("stats done", c1.max().item(), c1.min().item(), (s1 - s2).abs().max().item())


def stage_trace():
    """clock64 / globaltimer stamps of the pipeline roles of the TMA-fed pair kernel (developer build of the library)"""
    from vq_seg_b200 import _native, build
    build.build(dev=True)
    L = _native.use_dev_library()
    x, e = cases.FORWARD_CASES["c2_randn"]()
    xd, ed = x.to(dev), e.to(dev)
    xv = view(xd)
    blob = ops.prepare_codebook(ed)
    fused = "assign" not in sys.argv
    run = (lambda: ops._vq_forward_raw(xv, ed, blob, ops.MODE_TRAIN if "train" in sys.argv else ops.MODE_EVAL)) if fused else (lambda: ops.assign(xv, ed, blob, ops.ALGO_TC))
    for _ in range(3):
        run()
    buf = torch.zeros(148 * 4 * 256 + 8, dtype=torch.int64, device=dev)
    buf[-8] = 2 ** 62
    L.vqseg_debug_set_trace(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    L.vqseg_debug_set_trace(None)
    ex = buf[-8:].cpu()
    print('rescoring kernel: globaltimer span us', (ex[1] - ex[0]).item() / 1e3)
    t = buf[:-8].cpu().reshape(148, 4, 256)
    torch.save(t, os.path.join(ROOT, "gpurun_out", "trace.pt"))
    g = t[:, 3, 240:248]
    g0 = g[:, 0].min().item()
    for nm, col in (("entry", 0), ("setup done", 2), ("roles done (thread 0)", 4), ("exit", 6)):
        v = g[:, col] - g0
        print(f"globaltimer ns since first CTA entry: {nm:24s} min {v.min().item():7d} median {v.median().item():7d} max {v.max().item():7d}")
    for cta in (0, 1, 100, 147):
        tt = t[cta].clone()
        c0 = tt[3, 241].item()
        print(f"cta {cta}: clocks entry->setup {tt[3, 243].item() - c0}, ->roles done {tt[3, 245].item() - c0}, ->exit {tt[3, 247].item() - c0}")
        names = ["converter(w0)", "mma", "epilogue(w8)", "x loaders"]
        for role in range(4):
            row = tt[role, :240]
            ev = [(i, row[i].item() - c0) for i in range(240) if row[i] > 0]
            print(f"  {names[role]}: " + " ".join(f"{i}:{v}" for i, v in ev))


def stage_trace4():
    """clock64 stamps of the streaming pair kernel's roles (developer build): where a unit's time goes"""
    from vq_seg_b200 import _native, build
    build.build(dev=True)
    L = _native.use_dev_library()
    which = sys.argv[2] if len(sys.argv) > 2 else "c5"
    g = torch.Generator(device="cuda").manual_seed(3)
    if which == "c5":
        n, d, k = 148 * 128, 256, 8192          # one pair tile per CTA pair, 32 units
    else:
        n, d, k = 148 * 128 * 4, 512, 1024      # four pair tiles per CTA pair, 16 units
    xv = torch.randn(1, n, d, generator=g, device=dev)
    e = torch.randn(k, d, generator=g, device=dev)
    blob = ops.prepare_codebook(e)
    run = lambda: ops.assign(xv, e, blob, ops.ALGO_TC_STREAM_PAIR)
    for _ in range(3):
        run()
    buf = torch.zeros(148 * 4 * 256 + 8, dtype=torch.int64, device=dev)
    buf[-8] = 2 ** 62
    L.vqseg_debug_set_trace(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    L.vqseg_debug_set_trace(None)
    t = buf[:-8].cpu().reshape(148, 4, 256)
    n_dc = d // 64
    n_units = (k // 256) * (1 if which == "c5" else 4)
    for cta in (0, 1, 100):
        tt = t[cta]
        ep = tt[2]
        c0 = ep[0].item()
        print(f"cta {cta}: epilogue(w8) per unit: [start-wait, waited, processed]")
        line = []
        for u in range(min(n_units, 60)):
            line.append(f"{u}:{ep[4*u].item()-c0}/{ep[4*u+1].item()-ep[4*u].item()}/{ep[4*u+2].item()-ep[4*u+1].item()}")
        print("   " + " ".join(line))
        if cta % 2 == 0:
            mm = tt[1]
            line = []
            for u in range(min(n_units, 56)):
                line.append(f"{u}:{mm[128+2*u].item()-c0}/{mm[128+2*u+1].item()-mm[128+2*u].item()}")
            print("  mma tempty [start, waited]: " + " ".join(line))
            line = []
            for bq in range(64):
                if mm[2*bq] > 0:
                    line.append(f"{bq}:{mm[2*bq].item()-c0}/{mm[2*bq+1].item()-mm[2*bq].item()}")
            print("  mma bready (last 64 stages mod) [start, waited]: " + " ".join(line))
        cv = tt[0]
        line = []
        for q in range(min(120, 2 * n_dc * 4)):
            if cv[2*q] > 0:
                line.append(f"{q}:{cv[2*q].item()-c0}/{cv[2*q+1].item()-cv[2*q].item()}")
        print("  converter(w0) boxes [start, waited]: " + " ".join(line))


if __name__ == "__main__":
    t0 = time.time()
    {"exact": stage_exact, "tc": stage_tc, "ops": stage_ops, "time": stage_time, "prof": stage_prof, "trace": stage_trace, "bw": stage_bw, "bwbulk": stage_bwbulk, "seghead": stage_seghead, "big": stage_big, "shapes": stage_shapes, "stream": stage_stream, "trace4": stage_trace4, "null": stage_null, "stats": stage_stats}[sys.argv[1]]()
    print(f"stage {sys.argv[1]} done in {time.time() - t0:.1f}s")
