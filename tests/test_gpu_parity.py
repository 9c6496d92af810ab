"""GPU parity tests (run on a B200 with -m gpu): the CUDA path through the C ABI vs the committed golden
vectors (produced by the live reference) and vs the CPU oracle on the same seeded inputs.

Bars (north_star): bit-exact indices and code counts; quantize / loss / gradients within 1e-5 relative
(in fact quantize and the k-means means come out bit-exact, which the tests assert where it holds by
construction)."""
import itertools
import os

import pytest
import torch

import cases
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vq_seg_b200 import _native
    _native.lib()                       # fail loudly if the extension is missing on a GPU box
    return torch.device("cuda:0")


def view(x):
    b, c, h, w = x.shape
    return x.reshape(b, c, h * w).permute(0, 2, 1)


def near_tie_ok(x, e, idx_gpu, idx_ref):
    """Contract (ii) of SURVEY §7.4: any disagreement must be a row whose two candidates are within
    2 ulp in the reference's own fp32 distances."""
    bad = (idx_gpu != idx_ref).nonzero()
    if bad.numel() == 0:
        return True
    d = O.euclidean_dist(view(x), e)
    for b, p in bad.tolist():
        dg, dr = d[b, p, idx_gpu[b, p]], d[b, p, idx_ref[b, p]]
        if not (dg - dr).abs() <= 2 * torch.finfo(torch.float32).eps * dr.abs():
            return False
    return True


@pytest.mark.parametrize("algo", ["exact", "tc"])
@pytest.mark.parametrize("name", list(cases.FORWARD_CASES))
def test_assign_matches_golden(dev, golden, name, algo):
    from vq_seg_b200 import ops
    x, e = cases.FORWARD_CASES[name]()
    rec = golden["forward"][name]
    assert cases.sha(x) == rec["x_sha"]
    xd, ed = x.to(dev), e.to(dev)
    blob = ops.prepare_codebook(ed) if algo == "tc" else None
    idx, counts = ops.assign(view(xd), ed, blob, ops.ALGO_TC if algo == "tc" else ops.ALGO_EXACT)
    gold = rec["idx"].reshape(idx.shape).to(torch.int64)
    assert torch.equal(idx.cpu(), gold), f"{(idx.cpu() != gold).sum().item()} index mismatches vs the reference"
    assert torch.equal(counts.cpu(), rec["counts"])


@pytest.mark.parametrize("kernel", ["stream", "pair", "tma", "stream_pair"])
@pytest.mark.parametrize("name", ["c2_randn", "c2_relu", "d64", "odd_7x7", "k_not_tile", "dup_codes", "equidistant", "x_equals_code"])
def test_every_tensor_core_kernel_on_resident_shapes(dev, golden, name, kernel):
    """Shapes that fit the codebook-resident kernels, forced through each of the four tcgen05 filters in turn
    (single-CTA streaming, CTA pair fed through registers, CTA pair fed by TMA, streaming CTA pair fed by TMA): all
    must reproduce the reference."""
    from vq_seg_b200 import ops
    x, e = cases.FORWARD_CASES[name]()
    rec = golden["forward"][name]
    xd, ed = x.to(dev), e.to(dev)
    blob = ops.prepare_codebook(ed)
    algo = {"stream": ops.ALGO_TC_STREAM, "pair": ops.ALGO_TC_PAIR, "tma": ops.ALGO_TC_TMA,
            "stream_pair": ops.ALGO_TC_STREAM_PAIR}[kernel]
    if kernel in ("tma", "stream_pair") and (x.shape[2] * x.shape[3]) % 4 != 0:
        # a tensor map needs 16-byte strides: such maps are refused when forced (ALGO_TC / AUTO take the register-fed kernel)
        with pytest.raises(RuntimeError, match="not supported"):
            ops.assign(view(xd), ed, blob, algo)
        return
    idx, counts = ops.assign(view(xd), ed, blob, algo)
    torch.cuda.synchronize()
    assert torch.equal(idx.cpu(), rec["idx"].reshape(idx.shape).to(torch.int64))
    assert torch.equal(counts.cpu(), rec["counts"])


@pytest.mark.parametrize("shape", [(2, 256, 32, 32, 512), (2, 1024, 16, 16, 512), (1, 520, 12, 12, 300), (2, 64, 24, 24, 4096)])
@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_short_list_overflow_rows(dev, shape, metric):
    """More than 8 codes within the filter's error bound of a row's minimum (clusters of near-duplicate codes; a
    zero feature vector against a symmetric codebook): the rows are deferred to the block-per-row brute-force kernel.
    Indices and counts must equal the exact scorer's, including the first-index rule among exact duplicates."""
    from vq_seg_b200 import ops
    b, c, h, w, k = shape
    g = torch.Generator(device="cuda").manual_seed(c + k)
    x = torch.relu(torch.randn(b, c, h * w, generator=g, device=dev))
    xv = x.permute(0, 2, 1)
    base = xv.reshape(-1, c)[torch.randperm(b * h * w, generator=g, device=dev)[:k // 16]]
    e = base.repeat_interleave(16, dim=0)[:k].clone()                  # 16 copies of each centre ...
    e[1::2] += 1e-6 * torch.randn(e[1::2].shape, generator=g, device=dev)     # ... half of them perturbed in the last bits
    if e.shape[0] < k:
        e = torch.cat([e, torch.randn(k - e.shape[0], c, generator=g, device=dev)])
    e = e.contiguous()
    x[0, :, 5] = 0.0                                                   # a zero pixel
    ip = metric == "cosine"
    if ip:
        xv = ops.l2norm_rows(xv)
        ops.l2norm_rows_(e)
    flag = ops.METRIC_IP if ip else 0
    i_ex, c_ex = ops.assign(xv, e, None, ops.ALGO_EXACT | flag)
    i_tc, c_tc = ops.assign(xv, e, ops.prepare_codebook(e, ip), ops.ALGO_TC | flag)
    n_ovf = ops._last_assign_ws[4:8].view(torch.int32).item()
    assert n_ovf > 0, "the case is meant to overflow the short-list"
    assert torch.equal(i_ex, i_tc), f"{(i_ex != i_tc).sum().item()} rows differ ({n_ovf} overflow rows)"
    assert torch.equal(c_ex, c_tc)


@pytest.mark.parametrize("k,d", [(8192, 64), (65536, 32), (600, 256)])
def test_few_overflow_rows_are_split_over_blocks(dev, k, d):
    """A handful of overflow rows against many codes (zero feature vectors whose nearest codes are a dozen codes of
    equal norm): each row's code slices are spread over several blocks of the overflow kernel that meet in a packed
    atomicMin -- indices, counts and the sharded mode's (distance, index) keys must equal the exact scorer's."""
    from vq_seg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(k + d)
    x = torch.randn(1, d, 2048, generator=g, device=dev)
    x[0, :, [7, 900, 2047]] = 0.0
    xv = x.permute(0, 2, 1)
    e = torch.randn(k, d, generator=g, device=dev)
    # a dozen codes of one norm, just below every other code's: nearest to the zero vectors only
    e[100:112] = 0.8 * float(e.norm(dim=-1).min()) * torch.nn.functional.normalize(e[100:112], dim=-1)
    e[105] = e[101]                                                    # an exact duplicate: the lower index must win
    e = e.contiguous()
    blob = ops.prepare_codebook(e)
    i_ex, c_ex = ops.assign(xv, e, None, ops.ALGO_EXACT)
    i_tc, c_tc = ops.assign(xv, e, blob, ops.ALGO_AUTO)
    n_ovf = ops._last_assign_ws[4:8].view(torch.int32).item()
    assert 3 <= n_ovf <= 16, n_ovf
    assert torch.equal(i_ex, i_tc) and torch.equal(c_ex, c_tc)
    assert int(i_tc[0, 7]) in range(100, 112) and int(i_tc[0, 7]) != 105
    assert torch.equal(ops.assign_keys(xv, e, blob, 1000, ops.ALGO_AUTO), ops.assign_keys(xv, e, None, 1000, ops.ALGO_EXACT))


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("shape,layout", [((2, 1024, 32, 32, 512), "nchw"), ((2, 2048, 16, 16, 512), "nchw"),
                                          ((4, 1024, 32, 32, 512), "nchw"), ((1, 768, 20, 24, 700), "nchw"),
                                          ((1, 1536, 1, 1000, 300), "rows"), ((3, 640, 10, 12, 1024), "nchw")])
def test_split_d_mode_equals_exact(dev, shape, layout, metric):
    """Few rows, many dims (the model's 1024- / 2048-channel layers): the streaming pair kernel runs in split-D mode
    (partial scores summed by atomics in a scratch, short-lists built from the sums).  Indices, counts and the packed
    (distance, index) keys of the sharded mode must equal the exact scorer's, run after run."""
    from vq_seg_b200 import ops, _native
    b, c, h, w, k = shape
    g = torch.Generator(device="cuda").manual_seed(c + k)
    x = torch.relu(torch.randn(b, c, h * w, generator=g, device=dev))
    xv = x.permute(0, 2, 1)
    if layout == "rows":
        xv = xv.contiguous()
    e = (xv.reshape(-1, c)[torch.randint(0, b * h * w, (k,), generator=g, device=dev)] +
         0.2 * torch.randn(k, c, generator=g, device=dev)).contiguous()
    ip = metric == "cosine"
    if ip:
        xv = ops.l2norm_rows(xv)
        ops.l2norm_rows_(e)
    flag = ops.METRIC_IP if ip else 0
    blob = ops.prepare_codebook(e, ip)
    i_ex, c_ex = ops.assign(xv, e, None, ops.ALGO_EXACT | flag)
    for _ in range(3):
        i_tc, c_tc = ops.assign(xv, e, blob, ops.ALGO_AUTO | flag)
        assert torch.equal(i_ex, i_tc) and torch.equal(c_ex, c_tc)
    # the score scratch (last part of the workspace) was used: this really was the split-D path
    n, kp = b * h * w, (k + 255) // 256 * 256
    scratch = (n * kp * 4 + 255) // 256 * 256 + (n * 8 + 255) // 256 * 256
    ws = ops._last_assign_ws
    assert ws.numel() == _native.lib().vqseg_assign_workspace_bytes(n, c, k, 0)
    assert ws[ws.numel() - scratch:].view(torch.float32)[: n * kp].abs().sum().item() > 0
    if not ip:                                         # sharded mode: exact distances for every row
        keys_ex = ops.assign_keys(xv, e, None, 1000, ops.ALGO_EXACT)
        keys_tc = ops.assign_keys(xv, e, blob, 1000, ops.ALGO_AUTO)
        assert torch.equal(keys_ex, keys_tc)


@pytest.mark.parametrize("shape,layout", [((1, 70000, 512, 1024), "rows"), ((1, 33001, 256, 4096), "rows"),
                                          ((3, 9000, 320, 700), "rows"), ((8, 9216, 128, 2000), "nchw"), ((1, 300, 512, 600), "rows")])
def test_prepared_samples_equal_plain_assignment(dev, shape, layout):
    """vqseg_samples_prepare_f32 + vqseg_assign_prepared_f32 (the filter streams ready-made fp16 tiles: k-means assigns
    the same samples every Lloyd iteration): indices and counts identical to the exact scorer and to the plain path,
    for ragged row counts (odd tile counts, a padded last pair tile) and both layouts of the source."""
    from vq_seg_b200 import ops
    B, P, D, K = shape
    g = torch.Generator(device="cuda").manual_seed(B + P + D + K)
    centres = torch.randn(K, D, generator=g, device=dev)
    pick = torch.randint(0, K, (B, P), generator=g, device=dev)
    rows = centres[pick] * 0.7 + 0.8 * torch.randn(B, P, D, generator=g, device=dev)
    xv = rows.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "nchw" else rows
    blob = ops.prepare_codebook(centres)
    samples = ops.prepare_samples(xv)
    i_ex, c_ex = ops.assign(xv, centres, None, ops.ALGO_EXACT)
    i_pl, c_pl = ops.assign(xv, centres, blob, ops.ALGO_TC_STREAM_PAIR if P % 4 == 0 or layout == "rows" else ops.ALGO_TC)
    for _ in range(2):
        i_pr, c_pr = ops.assign(xv, centres, blob, ops.ALGO_AUTO, 0, samples)
        assert torch.equal(i_ex, i_pr), f"{(i_ex != i_pr).sum().item()} rows differ"
        assert torch.equal(c_ex, c_pr)
    assert torch.equal(i_pl, i_pr) and torch.equal(c_pl, c_pr)
    # new means, same samples (what a Lloyd iteration does)
    centres2 = centres + 0.05 * torch.randn(K, D, generator=g, device=dev)
    i2, c2 = ops.assign(xv, centres2, ops.prepare_codebook(centres2), ops.ALGO_AUTO, 0, samples)
    i2e, c2e = ops.assign(xv, centres2, None, ops.ALGO_EXACT)
    assert torch.equal(i2, i2e) and torch.equal(c2, c2e)


def test_kmeans_with_prepared_samples_is_bit_equal(dev):
    """kmeans() prepares large sample sets once; the result must equal the iteration-by-iteration conversion."""
    import vq_seg_b200 as V
    from vq_seg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(77)
    centres = torch.randn(600, 192, generator=g, device=dev)
    x = centres[torch.randint(0, 600, (1, 80000), generator=g, device=dev)] + 0.5 * torch.randn(1, 80000, 192, generator=g, device=dev)
    init = torch.randperm(80000, generator=g, device=dev)[:600]
    m1, b1 = V.kmeans(x, 600, 3, init_indices=init)                       # eligible: prepared samples
    m2, b2 = V.kmeans(x, 600, 3, init_indices=init, algo=ops.ALGO_TC_STREAM_PAIR)   # forced plain streaming kernel
    m3, b3 = V.kmeans(x, 600, 3, init_indices=init, algo=ops.ALGO_EXACT)
    assert torch.equal(b1, b2) and torch.equal(m1, m2) and torch.equal(b1, b3) and torch.equal(m1, m3)


def test_kblock_override_changes_only_near_ties(dev):
    """kblock is the fp32 chain split of the exact scorer; a different split may only move rows
    whose two best reference distances are within 2 ulp."""
    from vq_seg_b200 import ops
    x, e = cases.FORWARD_CASES["c1_l4"]()
    xd, ed = x.to(dev), e.to(dev)
    i_auto, _ = ops.assign(view(xd), ed, None, ops.ALGO_EXACT, 0)
    i_one, _ = ops.assign(view(xd), ed, None, ops.ALGO_EXACT, 1 << 20)       # one chain, no split
    assert near_tie_ok(x, e, i_one.cpu(), i_auto.cpu())


@pytest.mark.parametrize("layout", ["nchw", "rows"])
@pytest.mark.parametrize("shape", [
    # (B, P, D, K): the k-means shapes of BASELINE configs 4 / 5 in small, ragged tiles, few and many dim chunks
    (2, 4096, 512, 1024), (3, 1000, 512, 700), (1, 20000, 256, 4096), (2, 3000, 128, 2304), (4, 2500, 40, 3000),
    (1, 131, 512, 1024), (5, 36, 320, 1500), (2, 8192, 64, 256),
])
def test_streaming_pair_kernel_equals_exact(dev, shape, layout):
    """The TMA-fed streaming CTA-pair filter (assign_tc4.cu; k-means assignment, vq_img.py:39-41, and codebooks too
    large to stay resident) forced on NCHW maps and on packed sample rows: indices and counts bit-equal to the
    exact scorer (which the golden tests pin to the reference)."""
    from vq_seg_b200 import ops
    B, P, D, K = shape
    if layout == "nchw" and P % 4 != 0:
        P += 4 - P % 4                                   # tensor maps need 16-byte strides
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + D + K)
    centres = torch.randn(K, D, generator=g, device=dev)
    pick = torch.randint(0, K, (B, P), generator=g, device=dev)
    rows = centres[pick] * 0.7 + 0.8 * torch.randn(B, P, D, generator=g, device=dev)
    xv = rows.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "nchw" else rows
    blob = ops.prepare_codebook(centres)
    i1, c1 = ops.assign(xv, centres, None, ops.ALGO_EXACT)
    i2, c2 = ops.assign(xv, centres, blob, ops.ALGO_TC_STREAM_PAIR)
    torch.cuda.synchronize()
    assert torch.equal(i1, i2), f"{(i1 != i2).sum().item()} of {i1.numel()} rows differ"
    assert torch.equal(c1, c2)


def test_large_codebook_streaming(dev):
    """config-5-like: K = 8192 codes streamed through the single-CTA kernel (32 code chunks per tile)."""
    from vq_seg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(2, 128, 2048, generator=g, device=dev)
    e = torch.randn(8192, 128, generator=g, device=dev)
    xv = x.permute(0, 2, 1)
    i1, c1 = ops.assign(xv, e, None, ops.ALGO_EXACT)
    i2, c2 = ops.assign(xv, e, ops.prepare_codebook(e), ops.ALGO_TC)
    assert torch.equal(i1, i2) and torch.equal(c1, c2)
    assert near_tie_ok(x.cpu().reshape(2, 128, 2048, 1), e.cpu(), i2.cpu(), O.assign_euclidean(xv.cpu(), e.cpu()))


@pytest.mark.parametrize("name", ["c2_randn", "c2_relu", "c1_l3", "c1_l4", "c1_l5", "r448_l3", "r448_l5", "odd_7x7",
                                  "dup_codes", "x_equals_code", "equidistant", "all_zero_x", "n_lt_k", "k_not_tile",
                                  "d96_k1024", "c1_l5_uniform"])
def test_module_forward_backward_matches_golden(dev, golden, name):
    import vq_seg_b200 as V
    x, e = cases.FORWARD_CASES[name]()
    rec = golden["forward"][name]
    m = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0]).to(dev)
    m.codebook.embedding.weight.data.copy_(e)
    m.eval()
    with torch.no_grad():
        q, idx, loss, usage = m(x.to(dev))
    assert q.shape == x.shape and q.dtype == torch.float32 and q.is_contiguous()
    assert idx.shape == (x.shape[0], x.shape[2], x.shape[3]) and idx.dtype == torch.int64
    assert loss.shape == (1,) and not loss.requires_grad and usage.shape == () and usage.device == q.device
    assert torch.equal(idx.cpu().to(torch.int32), rec["idx"])
    assert cases.sha(q.cpu()) == rec["q_eval_sha"]                      # quantize == E[idx] bit-exact
    assert torch.equal(usage.cpu(), rec["usage"]) and torch.equal(loss.cpu(), rec["loss_eval"])
    m.train()
    xg = x.to(dev).requires_grad_(True)
    q, idx, loss, usage = m(xg)
    assert loss.requires_grad and loss.shape == (1,)
    assert cases.sha(q.detach().cpu()) == rec["q_train_sha"]            # x + (e - x), two roundings, bit-exact
    torch.testing.assert_close(loss.detach().cpu(), rec["loss_train"], rtol=1e-5, atol=0)
    g = torch.Generator().manual_seed(999)
    gq = torch.randn(q.shape, generator=g)
    (q * gq.to(dev)).sum().backward(retain_graph=True)
    assert torch.equal(xg.grad.cpu(), gq)                               # straight-through: identity
    xg.grad = None
    (loss * 1.5).sum().backward()
    step = max(1, x.numel() // 4096)
    torch.testing.assert_close(xg.grad.cpu().reshape(-1)[::step], rec["gx_l_sample"], rtol=1e-5, atol=1e-12)
    assert m.codebook.embedding.weight.grad is None                     # the codebook gets no gradient in training


@pytest.mark.parametrize("name", cases.KMEANS_CASES)
def test_kmeans_matches_golden(dev, golden, name):
    import vq_seg_b200 as V
    from vq_seg_b200 import ops
    x, k, iters, init_idx, use_cos = cases.kmeans_case(name)
    rec = golden["kmeans"][name]
    xv = view(x.to(dev))
    src = ops.l2norm_rows(xv) if use_cos else xv
    means, bins = V.kmeans(src, k, iters, use_cosine_sim=use_cos, init_indices=init_idx)
    assert means.shape == (1, k, x.shape[1]) and bins.shape == (1, k) and bins.dtype == torch.int64
    assert torch.equal(bins[0].cpu(), rec["bins"])                      # bit-exact counts after `iters` Lloyd steps
    assert torch.equal(means[0].cpu(), rec["means"])    # ordered per-code sums (+ ATen-order l2norm for cosine) -> bit-exact means


def test_kmeans_init_module_first_train_forward(dev):
    """kmeans_init=True: the first TRAINING forward runs k-means on that batch and overwrites the codebook
    (vq_img.py:165-166,179-190); eval forwards never do; `initted` is a plain attribute."""
    import vq_seg_b200 as V
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn(2, 64, 16, 16, generator=g))
    init = torch.randperm(512, generator=g)[:32]
    m = V.VectorQuantizer(dim=64, num_embeddings=32, kmeans_init=True).to(dev)
    w0 = m.codebook.embedding.weight.detach().clone()
    m.eval()
    with torch.no_grad():
        m(x.to(dev))
    assert not m.codebook.initted and torch.equal(m.codebook.embedding.weight, w0)
    m.train()
    m.codebook.kmeans_init_indices = init
    q, idx, loss, usage = m(x.to(dev))
    assert m.codebook.initted
    port = O.OracleVectorQuantizer(dim=64, num_embeddings=32, kmeans_init=True)
    port.kmeans_init_indices = init
    port.train()
    qo, io, lo, uo = port(x)
    assert torch.equal(m.codebook.embedding.weight.detach().cpu(), port.codebook.embedding.weight.detach())
    assert torch.equal(idx.cpu(), io) and torch.equal(q.detach().cpu(), qo.detach())
    assert "initted" not in m.state_dict()


@pytest.mark.parametrize("name", list(cases.COSINE_CASES))
def test_cosine_codebook(dev, golden, name):
    """round-1 goldens of the cosine codebook (vq_img.py:65-130), now bit-exact: indices, the straight-through output
    and the in-place renormalised weights."""
    import vq_seg_b200 as V
    x, e = cases.COSINE_CASES[name]()
    rec = golden["cosine"][name]
    m = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0], distance="cosine").to(dev)
    m.codebook.embedding.weight.data.copy_(e)
    m.train()
    q, idx, loss, usage = m(x.to(dev))
    assert torch.equal(idx.cpu(), rec["idx"].to(torch.int64))
    assert cases.sha(q) == rec["q_train_sha"]
    assert cases.sha(m.codebook.embedding.weight.data) == rec["weight_after_sha"]
    torch.testing.assert_close(usage.cpu(), rec["usage"])
    torch.testing.assert_close(loss.detach().cpu(), rec["loss_train"], rtol=1e-6, atol=0)


@pytest.fixture(scope="module")
def golden_cosine():
    return torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cosine_v2.pt"),
                      weights_only=False)["cosine2"]


def test_l2norm_matches_aten_summation_orders(dev):
    """vqseg_l2norm_f32 against F.normalize on the reference's CPU (SHA-256 goldens, make_golden_cosine.py): contiguous
    rows of every tail length D % 8, in place and out of place, and strided (B, HW, C) views with ragged pixel counts."""
    from vq_seg_b200 import ops
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cosine_v2.pt"),
                      weights_only=False)["l2norm"]
    g = torch.Generator().manual_seed(4242)
    for d in cases.L2NORM_LASTDIM_DIMS:
        e = torch.randn(200, d, generator=g)
        rec = gold[f"rows_d{d}"]
        assert cases.sha(e) == rec["in_sha"]
        out = ops.l2norm_rows(e.to(dev).unsqueeze(0))[0]
        assert cases.sha(out) == rec["out_sha"], f"D = {d}: contiguous rows"
        w = e.to(dev).clone()
        ops.l2norm_rows_(w)
        assert cases.sha(w) == rec["out_sha"], f"D = {d}: in place"
    for (b, c, p) in cases.L2NORM_VIEW_SHAPES:
        x = torch.randn(b, c, p, generator=g)
        rec = gold[f"view_{b}x{c}x{p}"]
        assert cases.sha(x) == rec["in_sha"]
        out = ops.l2norm_rows(x.to(dev).permute(0, 2, 1))
        assert cases.sha(out) == rec["out_sha"], f"view {(b, c, p)}"


@pytest.mark.parametrize("name", list(cases.COSINE2_CASES))
def test_cosine_codebook_bit_exact_train_and_eval(dev, golden_cosine, name):
    """The cosine codebook against the live reference's outputs (tests/golden/make_golden_cosine.py): l2norm of the
    strided view and of the weights bit-equal (SHA-256), indices equal in TRAIN and in EVAL mode (first index on
    ties, zero rows, D crossing MKL's K-blocking thresholds, D not a multiple of 8), the outputs and the twice
    renormalised weights bit-equal; and the tcgen05 path equal to the exact scorer."""
    import vq_seg_b200 as V
    from vq_seg_b200 import ops
    x, e = cases.COSINE2_CASES[name]()
    rec = golden_cosine[name]
    assert cases.sha(x) == rec["x_sha"] and cases.sha(e) == rec["e_sha"]
    xd, ed = x.to(dev), e.to(dev)
    xv = view(xd)
    xn = ops.l2norm_rows(xv)
    assert xn.stride() == xv.contiguous().permute(0, 2, 1).contiguous().permute(0, 2, 1).stride()    # layout kept
    assert cases.sha(xn) == rec["xn_sha"], "l2norm of the (B, HW, C) view differs from F.normalize on CPU"
    w1 = ed.clone()
    ops.l2norm_rows_(w1)
    assert cases.sha(w1) == rec["w1_sha"], "l2norm of the codebook rows differs from F.normalize on CPU"
    i_ex, c_ex = ops.assign_cosine(xn, w1, None, ops.ALGO_EXACT)
    i_tc, c_tc = ops.assign_cosine(xn, w1, None, ops.ALGO_TC)
    assert torch.equal(i_ex, i_tc) and torch.equal(c_ex, c_tc)
    assert torch.equal(i_ex.cpu().reshape(-1), rec["idx_train"].to(torch.int64).reshape(-1))

    m = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0], distance="cosine").to(dev)
    m.codebook.embedding.weight.data.copy_(ed)
    m.train()
    q, idx, loss, usage = m(xd)
    assert torch.equal(idx.cpu(), rec["idx_train"].to(torch.int64))
    assert cases.sha(q) == rec["q_train_sha"]
    assert cases.sha(m.codebook.embedding.weight.data) == rec["w_after_train_sha"]
    torch.testing.assert_close(usage.cpu(), rec["usage_train"])
    torch.testing.assert_close(loss.detach().cpu(), rec["loss_train"], rtol=1e-6, atol=0)
    m.eval()
    with torch.no_grad():
        q, idx, loss, usage = m(xd)
    assert torch.equal(idx.cpu(), rec["idx_eval"].to(torch.int64))
    assert cases.sha(q) == rec["q_eval_sha"]
    assert cases.sha(m.codebook.embedding.weight.data) == rec["w_after_eval_sha"]
    torch.testing.assert_close(usage.cpu(), rec["usage_eval"])
    assert loss.shape == (1,) and loss.item() == 0.0 and not loss.requires_grad


def test_kmeans_helpers_match_reference_semantics(dev):
    """batched_bincount (vq_img.py:22-27), sample_vectors both branches (:10-17) and batched_sample_vectors (:19-20)
    called directly, against the oracle / the reference's definitions."""
    import vq_seg_b200 as V
    g = torch.Generator().manual_seed(123)
    # batched_bincount == zeros.scatter_add_(-1, x, ones) == bincount with minlength, including empty bins
    for n, k in ((1000, 37), (5, 64), (100000, 512)):
        b = torch.randint(0, k, (1, n), generator=g)
        b[0, : n // 2] = b[0, : n // 2] % max(1, k // 3)                 # leave high bins sparse / empty
        got = V.batched_bincount(b.to(dev), k)
        want = torch.zeros(1, k, dtype=torch.int64).scatter_add_(-1, b, torch.ones_like(b))
        assert got.dtype == torch.int64 and got.shape == (1, k) and torch.equal(got.cpu(), want)
    # sample_vectors: injected indices reproduce sample[indices] bit for bit (both branches of :12-15)
    sample = torch.randn(300, 24, generator=g)
    for num in (64, 300, 1000):                                          # N > num, N == num, N < num (with replacement)
        ids = O.sample_indices(sample.shape[0], num, torch.Generator().manual_seed(num))
        assert ids.shape == (num,) and (len(set(ids.tolist())) == num) == (num <= 300)
        got = V.sample_vectors(sample.to(dev), num, ids.to(dev))
        assert torch.equal(got.cpu(), sample[ids])
    # ... and with the device RNG: rows of the sample, distinct when N >= num, any valid rows otherwise
    rows = {tuple(r.tolist()): i for i, r in enumerate(sample)}
    for num in (64, 1000):
        got = V.sample_vectors(sample.to(dev), num).cpu()
        assert got.shape == (num, 24)
        picked = [rows[tuple(r.tolist())] for r in got]                   # KeyError = not a row of the sample
        assert (len(set(picked)) == num) == (num <= 300)
    bs = V.batched_sample_vectors(sample.to(dev).unsqueeze(0), 50).cpu()
    assert bs.shape == (1, 50, 24) and all(tuple(r.tolist()) in rows for r in bs[0])
    # a strided NCHW view is sampled in place: the same rows as the contiguous copy
    x = torch.randn(2, 24, 6, 7, generator=g).to(dev)
    xv = view(x)
    ids = torch.tensor([0, 41, 42, 83, 7], device=dev)
    assert torch.equal(V.vq_img._sample_rows(xv, 5, ids), xv.reshape(-1, 24)[ids])


def test_module_cuda_graph_mode_and_weight_updates(dev, golden):
    """enable_cuda_graphs(): the no-grad eval forward replayed as one CUDA graph per input buffer gives the golden
    outputs, follows `weight.data` updates made behind the cache (the on-device codebook guard, ADVICE r1: the
    prepared blob is keyed on a version counter that .data writes do not bump), and leaves training forwards alone."""
    import vq_seg_b200 as V
    x, e = cases.FORWARD_CASES["c2_relu"]()
    rec = golden["forward"]["c2_relu"]
    m = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0]).to(dev)
    m.codebook.embedding.weight.data.copy_(e.to(dev))
    m.eval().enable_cuda_graphs()
    xd = x.to(dev)
    with torch.no_grad():
        for _ in range(3):                                   # capture, then two replays
            q, idx, loss, usage = m(xd)
    assert torch.equal(idx.cpu().to(torch.int32), rec["idx"]) and cases.sha(q) == rec["q_eval_sha"]
    assert loss.shape == (1,) and loss.item() == 0.0 and usage.item() == rec["usage"].item()
    assert len(m._graphs) == 1
    # the weights change through .data (k-means init, mean-teacher update): the replayed graph must see it
    g = torch.Generator().manual_seed(5)
    e2 = e + 0.5 * torch.randn(e.shape, generator=g)
    m.codebook.embedding.weight.data.copy_(e2.to(dev))
    with torch.no_grad():
        q2, idx2, _, _ = m(xd)
    want = O.assign_euclidean(view(x), e2)
    assert torch.equal(idx2.cpu().reshape(-1), want.reshape(-1)) or near_tie_ok(x, e2, idx2.cpu().reshape(want.shape), want)
    assert torch.equal(q2.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).cpu(), e2[idx2.cpu().reshape(-1)])
    # the same without graphs (plain forward after a .data write, no invalidate())
    m2 = V.VectorQuantizer(dim=x.shape[1], num_embeddings=e.shape[0]).to(dev).eval()
    m2.codebook.embedding.weight.data.copy_(e.to(dev))
    with torch.no_grad():
        m2(xd)
        m2.codebook.embedding.weight.data.copy_(e2.to(dev))
        _, idx3, _, _ = m2(xd)
    assert torch.equal(idx3, idx2)
    # a different buffer is a second entry; a training forward does not go through the cache
    with torch.no_grad():
        m(xd.clone())
    assert len(m._graphs) == 2
    m.train()
    q, idx, loss, usage = m(xd.clone().requires_grad_(True))
    assert loss.requires_grad and len(m._graphs) == 2


def test_row_major_samples_and_half_inputs(dev):
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 3000, 128, generator=g)            # packed (N, D) rows: sP = D, sD = 1
    e = torch.randn(200, 128, generator=g)
    ref = O.assign_euclidean(x, e)
    for algo in (ops.ALGO_EXACT, ops.ALGO_TC):
        blob = ops.prepare_codebook(e.to(dev))
        idx, counts = ops.assign(x.to(dev), e.to(dev), blob, algo)
        assert near_tie_ok(x.permute(0, 2, 1).reshape(1, 128, 3000, 1), e, idx.cpu(), ref)
        assert torch.equal(counts.cpu(), torch.bincount(idx.cpu().reshape(-1), minlength=200))
    import vq_seg_b200 as V
    xh = torch.randn(2, 64, 12, 12, generator=g).half()
    m = V.VectorQuantizer(dim=64, num_embeddings=40).to(dev)
    eh = torch.randn(40, 64, generator=g)
    m.codebook.embedding.weight.data.copy_(eh)
    m.train()
    q, idx, loss, _ = m(xh.to(dev))
    r = O.vq_forward(xh, eh, True, 1)                      # the reference upcasts with x.to(float32) (vq_img.py:229)
    assert q.dtype == torch.float32 and torch.equal(idx.cpu(), r["embed_index"]) and torch.equal(q.detach().cpu(), r["quantize"])


def test_large_random_sweep_tc_equals_exact(dev):
    """Size-independent property at BASELINE sizes: the tcgen05 filter + rescoring must equal the exact
    scorer bit for bit, whatever the data (that is the proof obligation of the short-list bound)."""
    from vq_seg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    for (b, c, hw, k, kind) in [(8, 256, 4096, 512, "randn"), (8, 256, 4096, 512, "relu"), (4, 512, 3136, 512, "relu"),
                                (1, 64, 20000, 1000, "randn"), (2, 2048, 256, 512, "relu"), (3, 40, 777, 33, "randn"),
                                (8, 256, 4096, 512, "clustered"), (2, 256, 4096, 512, "tiny"), (2, 256, 4096, 512, "huge")]:
        x = torch.randn(b, c, hw, generator=g, device=dev)
        e = torch.randn(k, c, generator=g, device=dev)
        if kind == "relu":
            x = torch.relu(x)
            e = x.permute(0, 2, 1).reshape(-1, c)[torch.randperm(b * hw, device=dev, generator=g)[:k]] \
                + 0.05 * torch.randn(k, c, generator=g, device=dev)
        elif kind == "clustered":      # many near-ties: codes are tight perturbations of 16 centres
            centres = torch.randn(16, c, generator=g, device=dev)
            e = centres[torch.arange(k, device=dev) % 16] + 1e-3 * torch.randn(k, c, generator=g, device=dev)
            x = centres[torch.randint(0, 16, (b * hw,), device=dev, generator=g)].reshape(b, hw, c).permute(0, 2, 1) \
                + 0.1 * torch.randn(b, c, hw, generator=g, device=dev)
        elif kind == "tiny":
            x, e = x * 1e-4, e * 1e-4
        elif kind == "huge":
            x, e = x * 3e3, e * 3e3
        xv = x.permute(0, 2, 1)
        blob = ops.prepare_codebook(e)
        i1, c1 = ops.assign(xv, e, None, ops.ALGO_EXACT)
        i2, c2 = ops.assign(xv, e, blob, ops.ALGO_TC)
        assert torch.equal(i1, i2), (kind, (i1 != i2).sum().item())
        assert torch.equal(c1, c2) and c1.sum().item() == b * hw
        # live oracle on the host CPU: bit-exact except provable fp32 near-ties
        ref = O.assign_euclidean(xv.cpu(), e.cpu())
        assert near_tie_ok(x.cpu().reshape(b, c, hw, 1), e.cpu(), i2.cpu(), ref), kind


def test_fp16_overflow_rows_fall_back_to_exact(dev):
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(13)
    x = torch.randn(1, 64, 512, generator=g)
    e = torch.randn(64, 64, generator=g)
    x[0, :, 5] *= 1e7                                   # overflows the fp16 operand of the filter
    x[0, 3, 100] = 3e6
    xv = x.permute(0, 2, 1).to(dev)
    i1, _ = ops.assign(xv, e.to(dev), None, ops.ALGO_EXACT)
    i2, _ = ops.assign(xv, e.to(dev), ops.prepare_codebook(e.to(dev)), ops.ALGO_TC)
    assert torch.equal(i1, i2)
    assert torch.equal(i2.cpu(), O.assign_euclidean(xv.cpu(), e))


def test_code_stats_paths(dev):
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(17)
    x = torch.randn(2, 96, 900, generator=g)
    idx = torch.randint(0, 50, (2, 900), generator=g)
    xv = x.permute(0, 2, 1)
    flat = xv.reshape(-1, 96)
    ref_sums = torch.zeros(50, 96).scatter_add_(0, idx.reshape(-1, 1).expand(-1, 96).contiguous(), flat.contiguous())
    ref_counts = torch.bincount(idx.reshape(-1), minlength=50)
    c, s = ops.code_stats(xv.to(dev), idx.to(dev), 50, True)
    assert torch.equal(c.cpu(), ref_counts) and torch.equal(s.cpu(), ref_sums)       # ordered sums: bit-exact
    c, s = ops.code_stats(xv.to(dev), idx.to(dev), 50, False)
    assert torch.equal(c.cpu(), ref_counts)
    torch.testing.assert_close(s.cpu(), ref_sums, rtol=1e-5, atol=1e-5)               # atomics: order not fixed
    c, s = ops.code_stats(flat.contiguous().unsqueeze(0).to(dev), idx.reshape(1, -1).to(dev), 50, True)
    assert torch.equal(s.cpu(), ref_sums)
    # more codes than the counter table of the fast scatter holds (the turn-taking scatter), more than one scan tile
    wide = torch.randint(0, 5000, (2, 900), generator=g)
    ref_wide = torch.zeros(5000, 96).scatter_add_(0, wide.reshape(-1, 1).expand(-1, 96).contiguous(), flat.contiguous())
    c, s = ops.code_stats(xv.to(dev), wide.to(dev), 5000, True)
    assert torch.equal(c.cpu(), torch.bincount(wide.reshape(-1), minlength=5000)) and torch.equal(s.cpu(), ref_wide)
    # one code owning 70 % of the pixels (the big-cluster kernel) on the NCHW view and on packed rows
    hot = torch.where(torch.rand(2, 900, generator=g) < 0.7, torch.zeros(2, 900, dtype=torch.long), idx)
    assert int((hot == 0).sum()) > 1024
    ref_hot = torch.zeros(50, 96).scatter_add_(0, hot.reshape(-1, 1).expand(-1, 96).contiguous(), flat.contiguous())
    for view in (xv, flat.contiguous().unsqueeze(0)):
        c, s = ops.code_stats(view.to(dev), hot.reshape(view.shape[0], -1).to(dev), 50, True)
        assert torch.equal(c.cpu(), torch.bincount(hot.reshape(-1), minlength=50)) and torch.equal(s.cpu(), ref_hot)


def test_code_stats_chunked_large(dev):
    """> 64 MB of rows: the ordered sums run range by range (TLB reach) and must still equal the sequential
    CPU scatter_add_ bit for bit; skewed clusters on purpose."""
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(23)
    n, d, k = 90_000, 512, 64
    x = torch.randn(n, d, generator=g)
    idx = (torch.rand(n, generator=g) ** 3 * k).long().clamp_(0, k - 1)          # heavy skew towards code 0
    ref_sums = torch.zeros(k, d).scatter_add_(0, idx.reshape(-1, 1).expand(-1, d).contiguous(), x)
    ref_counts = torch.bincount(idx, minlength=k)
    c, s = ops.code_stats(x.unsqueeze(0).to(dev), idx.unsqueeze(0).to(dev), k, True)
    assert torch.equal(c.cpu(), ref_counts) and torch.equal(s.cpu(), ref_sums)
    c, s = ops.code_stats(x.unsqueeze(0).to(dev), idx.unsqueeze(0).to(dev), k, False)
    assert torch.equal(c.cpu(), ref_counts)
    torch.testing.assert_close(s.cpu(), ref_sums, rtol=2e-5, atol=2e-3)
    # a big cluster (more than twice the mean share and > 4096 rows: the cp.async ring kernel) at a width that fills
    # neither a 32-dim warp nor a 128-dim slab
    n2, d2, k2 = 30_000, 100, 3
    x2 = torch.randn(n2, d2, generator=g)
    i2 = (torch.rand(n2, generator=g) ** 4 * k2).long().clamp_(0, k2 - 1)
    assert int(torch.bincount(i2, minlength=k2).max()) > 2 * n2 // k2
    ref2 = torch.zeros(k2, d2).scatter_add_(0, i2.reshape(-1, 1).expand(-1, d2).contiguous(), x2)
    c, s = ops.code_stats(x2.unsqueeze(0).to(dev), i2.unsqueeze(0).to(dev), k2, True)
    assert torch.equal(c.cpu(), torch.bincount(i2, minlength=k2)) and torch.equal(s.cpu(), ref2)
    # sorted assignments: the big code fills the first row ranges and is absent from the last one (empty segments on
    # the big-cluster path), the other codes live only in the later ranges (empty leading segments on the plain path)
    i3 = torch.cat([torch.zeros(54_000, dtype=torch.long), torch.randint(1, 8, (n - 54_000,), generator=g).sort().values])
    ref3 = torch.zeros(8, d).scatter_add_(0, i3.reshape(-1, 1).expand(-1, d).contiguous(), x)
    c, s = ops.code_stats(x.unsqueeze(0).to(dev), i3.unsqueeze(0).to(dev), 8, True)
    assert torch.equal(c.cpu(), torch.bincount(i3, minlength=8)) and torch.equal(s.cpu(), ref3)
    # >= 2^18 packed rows, unordered: the sorted-window kernel (one RED per code met in a window of 64 sorted rows);
    # a hot code, a width that is not a multiple of 128, some rows without a code
    n4, d4, k4 = 300_000, 200, 100
    x4 = torch.randn(n4, d4, generator=g)
    i4 = (torch.rand(n4, generator=g) ** 3 * k4).long().clamp_(0, k4 - 1)
    i4[::1000] = -1
    keep = i4 >= 0
    ref4 = torch.zeros(k4, d4).scatter_add_(0, i4[keep].reshape(-1, 1).expand(-1, d4).contiguous(), x4[keep])
    c, s = ops.code_stats(x4.unsqueeze(0).to(dev), i4.unsqueeze(0).to(dev), k4, False)
    assert torch.equal(c.cpu(), torch.bincount(i4[keep], minlength=k4))
    torch.testing.assert_close(s.cpu(), ref4, rtol=2e-5, atol=5e-3)
    c, s = ops.code_stats(x4.unsqueeze(0).to(dev), i4.unsqueeze(0).to(dev), k4, True)
    assert torch.equal(c.cpu(), torch.bincount(i4[keep], minlength=k4)) and torch.equal(s.cpu(), ref4)
    # many images (NCHW): whole-image groups per pass
    xi = torch.randn(40, 256, 4096, generator=g)                                  # 40 images x 4 MiB
    ii = torch.randint(0, k, (40, 4096), generator=g)
    flat = xi.permute(0, 2, 1).reshape(-1, 256)
    ref = torch.zeros(k, 256).scatter_add_(0, ii.reshape(-1, 1).expand(-1, 256).contiguous(), flat.contiguous())
    c, s = ops.code_stats(xi.permute(0, 2, 1).to(dev), ii.to(dev), k, True)
    assert torch.equal(s.cpu(), ref)


def test_amp_compat_rounds_code_through_fp16(dev):
    import vq_seg_b200 as V
    g = torch.Generator().manual_seed(19)
    x = torch.randn(2, 64, 8, 8, generator=g)
    e = torch.randn(32, 64, generator=g)
    m = V.VectorQuantizer(dim=64, num_embeddings=32).to(dev)
    m.codebook.embedding.weight.data.copy_(e)
    m.train()
    with torch.autocast("cuda", dtype=torch.float16):
        q, idx, loss, _ = m(x.to(dev))
    r = O.vq_forward(x, e, False)
    xb = view(x)
    expect = xb + (e[r["embed_index"].reshape(2, -1)].half().float() - xb)           # SURVEY §5 AMP policy
    assert q.dtype == torch.float32 and loss.dtype == torch.float32
    assert torch.equal(q.detach().cpu(), expect.permute(0, 2, 1).reshape(x.shape))
    assert torch.equal(idx.cpu(), r["embed_index"])


def test_eval_mode_gradient_goes_to_codebook(dev):
    import vq_seg_b200 as V
    m = V.VectorQuantizer(dim=16, num_embeddings=8).to(dev)
    m.eval()
    x = torch.randn(1, 16, 5, 5, device=dev, requires_grad=True)
    q, idx, _, _ = m(x)
    assert q.requires_grad
    q.sum().backward()
    assert x.grad is None
    expect = torch.bincount(idx.reshape(-1), minlength=8).float().unsqueeze(1).expand(8, 16)
    torch.testing.assert_close(m.codebook.embedding.weight.grad, expect)


def test_training_step_like_reference_loop(dev):
    """The VQ call pattern of the reference's model + training loop (modified_vqunet/net.py:226-237,
    train_vqreptunet1x1v2.py:143-202): make_vq_module list with Identity levels, k-means init on the first
    training forward, eval-mode no_grad pass, fp16 autocast + GradScaler, commitment loss summed over levels,
    `code_usage.detach().cpu()` per level, Adam over all parameters (codebook grads stay None)."""
    import vq_seg_b200 as V

    class TinyVQNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = torch.nn.ModuleList([torch.nn.Conv2d(3, 16, 3, 2, 1), torch.nn.Conv2d(16, 64, 3, 2, 1),
                                            torch.nn.Conv2d(64, 128, 3, 2, 1)])
            self.codebook = V.make_vq_module({"num_embeddings": [0, 64, 64], "distance": "euclidean", "kmeans_init": True},
                                             [3, 16, 64, 128], 3)
            self.head = torch.nn.Conv2d(128, 3, 1)

        def forward(self, x):
            feats = []
            for conv in self.enc:
                x = torch.relu(conv(x))
                feats.append(x)
            loss = torch.tensor([0.], device=x.device, requires_grad=self.training)
            usage = []
            for i in range(len(feats)):
                q, _idx, commit, cu = self.codebook[i](feats[i])
                feats[i] = q
                if commit is not None:
                    loss = loss + commit
                if cu is not None:
                    usage.append(cu.detach().cpu())
            loss = loss / len(feats)
            return self.head(feats[-1]), loss, torch.tensor(usage)

    torch.manual_seed(0)
    net = TinyVQNet().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    scaler = torch.amp.GradScaler("cuda", enabled=True)
    x = torch.rand(4, 3, 64, 64, device=dev)
    y = torch.randint(0, 3, (4, 8, 8), device=dev)
    net.eval()
    with torch.no_grad():
        net(x)                                            # eval pass first: must NOT trigger k-means
    assert not net.codebook[1].codebook.initted
    net.train()
    w_before = [m.codebook.embedding.weight.detach().clone() for m in net.codebook[1:]]
    losses = []
    for step in range(3):
        with torch.autocast("cuda", dtype=torch.float16):
            logits, commit, usage = net(x)
            loss = torch.nn.functional.cross_entropy(logits.float(), y) + commit.sum()
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(loss.item())
        assert usage.shape == (2,) and commit.shape == (1,) and commit.requires_grad
        if step == 0:
            w_init = [m.codebook.embedding.weight.detach().clone() for m in net.codebook[1:]]
            assert all(m.codebook.initted for m in net.codebook[1:])
            assert all(not torch.equal(a, b) for a, b in zip(w_before, w_init))     # k-means overwrote the N(0,1) init
    # the codebook is frozen after the k-means init: no gradient, Adam skips it
    assert all(m.codebook.embedding.weight.grad is None for m in net.codebook[1:])
    assert all(torch.equal(m.codebook.embedding.weight.detach(), w) for m, w in zip(net.codebook[1:], w_init))
    assert all(p.grad is not None for p in net.enc.parameters())                     # STE + commitment reach the encoder
    assert all(torch.isfinite(torch.tensor(losses)))


def test_opcheck(dev):
    from vq_seg_b200 import ops
    x = torch.randn(2, 16, 50, device=dev).permute(0, 2, 1)
    e = torch.randn(24, 16, device=dev)
    torch.library.opcheck(torch.ops.vqseg.assign, (x, e, None, 1, 0), test_utils=("test_schema", "test_faketensor"))
    idx, _ = ops.assign(x, e, None, 1)
    torch.library.opcheck(torch.ops.vqseg.gather_ste, (x, e, idx, 1), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(torch.ops.vqseg.code_stats, (x, idx, 24, True), test_utils=("test_schema", "test_faketensor"))


def test_ema_extension_matches_standard_equations(dev):
    """Opt-in EMA codebook update (no reference counterpart, parity unpinned): kernel against the restated standard
    VQ-VAE rule, then through the module: one training forward moves the codebook as the rule predicts and leaves
    the state_dict keys alone."""
    import vq_seg_b200 as V
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(21)
    k, d = 300, 96
    counts = torch.randint(0, 50, (k,), generator=g)
    counts[::7] = 0
    sums = torch.randn(k, d, generator=g) * counts.unsqueeze(1).float()
    cs0 = torch.rand(k, generator=g) * 20
    ea0 = torch.randn(k, d, generator=g)
    cs_ref, ea_ref, w_ref = O.ema_update(counts, sums, cs0, ea0, 0.8, 1e-5)
    cs, ea, w = cs0.to(dev), ea0.to(dev), torch.empty(k, d, device=dev)
    ops.ema_update(counts.to(dev), sums.to(dev), cs, ea, w, 0.8, 1e-5)
    # fp32 throughout; sums of opposite sign cancel in embed_avg, hence the absolute term
    assert torch.allclose(cs.cpu(), cs_ref, rtol=1e-6, atol=1e-6) and torch.allclose(ea.cpu(), ea_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(w.cpu(), w_ref, rtol=1e-5, atol=1e-6)
    x, e = cases.FORWARD_CASES["d64"]()
    m = V.VectorQuantizer(dim=64, num_embeddings=e.shape[0], decay=0.9).to(dev)
    m.codebook.embedding.weight.data.copy_(e.to(dev))
    m.enable_ema()
    assert list(m.state_dict()) == ["codebook.embedding.weight"]
    m.train()
    q, idx, loss, usage = m(x.to(dev))
    xv = view(x)
    cnt = torch.bincount(idx.cpu().reshape(-1), minlength=e.shape[0])
    sm = torch.zeros_like(e).index_add_(0, idx.cpu().reshape(-1), xv.reshape(-1, 64))
    _, _, w_exp = O.ema_update(cnt, sm, torch.ones(e.shape[0]), e.clone(), 0.9, 1e-5)       # cluster_size starts at one per code
    assert torch.equal(idx.cpu().reshape(xv.shape[:2]), O.assign_euclidean(xv, e))     # the lookup used the OLD codebook
    assert torch.allclose(m.codebook.embedding.weight.detach().cpu(), w_exp, rtol=1e-4, atol=1e-5)
    w1 = m.codebook.embedding.weight.detach().cpu().clone()
    q2, idx2, _, _ = m(x.to(dev))                                              # and the next forward the NEW one
    assert near_tie_ok(x, w1, idx2.cpu().reshape(xv.shape[:2]), O.assign_euclidean(xv, w1))
    m.eval()
    w_before = m.codebook.embedding.weight.detach().clone()
    m(x.to(dev))
    assert torch.equal(m.codebook.embedding.weight.detach(), w_before)         # no update outside training


def test_baseline_config_sizes_properties(dev):
    """BASELINE configs 4 and 5 at (or near) their full sizes, through properties that need no CPU pass:
    config 5's codebook (K = 65536, D = 256): filter + rescoring == brute-force exact scorer on 16 Ki rows;
    config 4's k-means (K = 1024, D = 512) on 2 Mi rows: counts add up, the ordered statistics are reproducible
    bit for bit, the atomic ones agree to 1e-5, assignment is a pure function of (x, means)."""
    from vq_seg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(41)
    n, d, k = 16384, 256, 65536
    x = torch.randn(1, n, d, generator=g, device=dev)
    e = torch.randn(k, d, generator=g, device=dev)
    i_tc, c_tc = ops.assign(x, e, ops.prepare_codebook(e), ops.ALGO_TC)
    i_ex, c_ex = ops.assign(x, e, None, ops.ALGO_EXACT)
    assert torch.equal(i_tc, i_ex) and torch.equal(c_tc, c_ex) and c_tc.sum().item() == n
    del x, e
    n, d, k = 1 << 21, 512, 1024
    x = torch.randn(1, n, d, generator=g, device=dev)
    means = x[0, torch.randperm(n, device=dev, generator=g)[:k]].contiguous()
    blob = ops.prepare_codebook(means)
    idx, counts = ops.assign(x, means, blob, ops.ALGO_TC)
    idx2, counts2 = ops.assign(x, means, blob, ops.ALGO_TC)
    assert torch.equal(idx, idx2) and torch.equal(counts, counts2) and counts.sum().item() == n
    b1, s1 = ops.code_stats(x, idx, k, True)
    b2, s2 = ops.code_stats(x, idx, k, True)
    b3, s3 = ops.code_stats(x, idx, k, False)
    assert torch.equal(b1, counts) and torch.equal(b3, counts)
    assert torch.equal(s1, s2)                                                  # deterministic order
    assert (s1 - s3).abs().max().item() <= 1e-5 * s1.abs().max().item()
    # the k-means start rows are samples: each is its own nearest code (distance exactly 0 up to fp32 cancellation)
    sub = x[:, :4096]
    assert torch.equal(ops.assign(sub, means, None, ops.ALGO_EXACT)[0], idx[:, :4096])


def test_random_shape_fuzz(dev):
    """40 random small problems (odd dims, single pixels, K from 1 to 700, row-major and NCHW views, train and eval):
    the fused forward must agree with the brute-force exact scorer + a torch gather on the same device, and with the
    CPU oracle's indices up to provable near-ties."""
    from vq_seg_b200 import ops
    rng = torch.Generator().manual_seed(1234)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))     # noqa: E731
    for it in range(40):
        b, c, hw, k = ri(1, 3), ri(1, 300), ri(1, 700), ri(1, 700)
        if it % 5 == 0:
            hw = 128 * ri(1, 6)
        x = torch.randn(b, c, hw, generator=rng)
        e = torch.randn(k, c, generator=rng)
        if it % 3 == 0:
            x = torch.relu(x)
        xd, ed = x.to(dev), e.to(dev)
        xv = xd.permute(0, 2, 1)
        if it % 4 == 1:
            xv = xv.contiguous()                                               # row-major samples
        blob = ops.prepare_codebook(ed)
        i_ex, c_ex = ops.assign(xv, ed, None, ops.ALGO_EXACT)
        for mode in (ops.MODE_TRAIN, ops.MODE_EVAL):
            q, idx, mse, usage = ops.vq_forward(xv, ed, blob, mode, ops.ALGO_AUTO)
            tag = (it, b, c, hw, k, mode)
            assert torch.equal(idx, i_ex), tag
            eq = ed[idx]
            want = xv + (eq - xv) if mode == ops.MODE_TRAIN else eq
            assert torch.equal(q, want), tag
            assert abs(usage.item() - 100.0 * (c_ex == 0).sum().item() / k) < 1e-4, tag
            if mode == ops.MODE_TRAIN:
                ref = ((want - xv) ** 2).double().mean().item()
                assert abs(mse.item() - ref) <= 1e-5 * max(ref, 1e-30), tag
        ref_idx = O.assign_euclidean(xv.cpu(), e)
        assert near_tie_ok(x.reshape(b, c, hw, 1), e, i_ex.cpu(), ref_idx), (it, b, c, hw, k)


def test_training_step_is_graph_capturable(dev):
    """The training forward + backward of the module under torch.cuda.make_graphed_callables (nothing on the path
    allocates through the driver, synchronises or reads back): outputs and the input gradient bit-equal to eager."""
    import vq_seg_b200 as V
    torch.manual_seed(11)
    m = V.VectorQuantizer(dim=96, num_embeddings=64).to(dev).train()
    m.codebook.embedding.weight.data.normal_()
    x = torch.randn(2, 96, 24, 24, device=dev).requires_grad_(True)
    gy = torch.randn(2, 96, 24, 24, device=dev)
    one = torch.ones(1, device=dev)

    def step(fn):
        q, idx, loss, usage = fn(x)
        torch.autograd.backward((q, loss), (gy, one))
        g = x.grad.clone()
        x.grad = None
        return q.detach().clone(), idx.clone(), loss.detach().clone(), usage.clone(), g

    eager = step(m)
    gm = torch.cuda.make_graphed_callables(m, (x.detach().clone().requires_grad_(True),), allow_unused_input=True)
    for _ in range(2):                                      # replays
        graphed = step(gm)
        for a, b in zip(eager, graphed):
            assert torch.equal(a, b)
    m.codebook.embedding.weight.data.mul_(0.5)              # the captured graph follows weight updates (codebook guard)
    e2, g2 = step(m), step(gm)
    assert not torch.equal(e2[0], eager[0])
    for a, b in zip(e2, g2):
        assert torch.equal(a, b)


def test_filter_slack_over_operand_scales(dev):
    """The filters' slack (common.cuh filter_slack) must stay a rigorous bound at every ratio of |e| to |x|: the
    reference's default uniform(-1/K, 1/K) init against unit-scale features (|e| << |x|, where the slack is governed
    by the few roundings at the magnitude of |x|^2), the opposite ratio, and equal scales -- every tensor-core filter
    that takes the shape against the brute-force exact scorer, and the default-init case against the CPU reference's
    cdist + argmin itself."""
    from vq_seg_b200 import ops
    g = torch.Generator().manual_seed(97)
    algos = (ops.ALGO_AUTO, ops.ALGO_TC_STREAM, ops.ALGO_TC_PAIR, ops.ALGO_TC_TMA, ops.ALGO_TC_STREAM_PAIR)
    undecided = {}
    for (d, k, p), xs, es, relu in itertools.product(((256, 512, 2048), (64, 200, 1024), (512, 1024, 1024), (1024, 512, 512)),
                                                      (1e-3, 1.0, 40.0), (1e-4, None, 1.0, 40.0), (False, True)):
        es = 1.0 / k if es is None else es
        x = torch.randn(2, d, p, generator=g) * xs
        if relu:
            x = torch.relu(x)
        e = (torch.rand(k, d, generator=g) * 2 - 1) * es
        xv, ed = x.to(dev).permute(0, 2, 1), e.to(dev).contiguous()
        blob = ops.prepare_codebook(ed)
        i_ex, c_ex = ops.assign(xv, ed, None, ops.ALGO_EXACT)
        for algo in algos:
            try:
                idx, counts = ops.assign(xv, ed, blob, algo)
            except RuntimeError as exc:
                assert "not supported" in str(exc), exc
                continue
            assert torch.equal(idx, i_ex) and torch.equal(counts, c_ex), (d, k, p, xs, es, relu, algo)
            if algo == ops.ALGO_AUTO:
                undecided[(d, xs, es, relu)] = int(ops._last_assign_ws[:4].view(torch.int32).item()) / (2 * p)
        if xs == 1.0 and es == 1.0 / k and d <= 256:
            assert torch.equal(i_ex.cpu(), O.assign_euclidean(x.permute(0, 2, 1), e))
    # the default init must not be a cliff: most rows are decided by the filter itself
    assert undecided[(256, 1.0, 1.0 / 512, True)] < 0.25, undecided[(256, 1.0, 1.0 / 512, True)]


def test_random_shape_fuzz_every_filter_and_metric(dev):
    """30 random problems through EVERY tensor-core filter that accepts the shape (single-CTA streaming, register-fed
    pair, TMA-fed pair, TMA-fed streaming pair) in both metrics: each must give the exact scorer's indices and counts;
    the cosine indices must equal the CPU reference's (einsum + argmax on F.normalize'd inputs) up to near-ties."""
    from vq_seg_b200 import ops
    rng = torch.Generator().manual_seed(4321)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))     # noqa: E731
    algos = {"stream": ops.ALGO_TC_STREAM, "pair": ops.ALGO_TC_PAIR, "tma": ops.ALGO_TC_TMA, "stream_pair": ops.ALGO_TC_STREAM_PAIR}
    ran = {k: 0 for k in algos}
    for it in range(30):
        b, c, hw, k = ri(1, 3), 4 * ri(2, 140), 4 * ri(8, 700), ri(30, 3000)
        x = torch.randn(b, c, hw, generator=rng)
        if it % 3 == 0:
            x = torch.relu(x)
        e = x.permute(0, 2, 1).reshape(-1, c)[torch.randint(0, b * hw, (k,), generator=rng)] + 0.3 * torch.randn(k, c, generator=rng)
        xd, ed = x.to(dev), e.to(dev).contiguous()
        xv = xd.permute(0, 2, 1)
        if it % 4 == 1:
            xv = xv.contiguous()
        for ip in (False, True):
            xx, ee = xv, ed
            if ip:
                xx = ops.l2norm_rows(xv)
                ee = ed.clone()
                ops.l2norm_rows_(ee)
            flag = ops.METRIC_IP if ip else 0
            blob = ops.prepare_codebook(ee, ip)
            i_ex, c_ex = ops.assign(xx, ee, None, ops.ALGO_EXACT | flag)
            for name, algo in algos.items():
                try:
                    idx, counts = ops.assign(xx, ee, blob, algo | flag)
                except RuntimeError as exc:
                    assert "not supported" in str(exc), exc
                    continue
                ran[name] += 1
                assert torch.equal(idx, i_ex) and torch.equal(counts, c_ex), (it, name, ip, b, c, hw, k)
            if ip:
                xn_ref = torch.nn.functional.normalize(xv.cpu(), p=2, dim=-1)
                en_ref = torch.nn.functional.normalize(e, p=2, dim=-1)
                sim = torch.einsum("n d, e d -> n e", xn_ref.contiguous().view(-1, c), en_ref).view(b, hw, k)
                ref = sim.argmax(-1)
                bad = (ref != i_ex.cpu()).nonzero()
                for bb, pp in bad.tolist():                      # MKL splits tiny problems along K: last-bit ties only
                    a, r = sim[bb, pp, i_ex[bb, pp].item()], sim[bb, pp, ref[bb, pp]]
                    assert (a - r).abs() <= 4 * torch.finfo(torch.float32).eps * r.abs().clamp_min(1e-30), (it, bb, pp)
    assert all(v > 0 for v in ran.values()), ran


def test_empty_batch(dev):
    """A batch of zero images: empty outputs, zero loss, every code unused (usage 100 %), no kernel fault."""
    import vq_seg_b200 as V
    m = V.VectorQuantizer(dim=64, num_embeddings=128).to(dev)
    for train in (True, False):
        m.train(train)
        q, idx, loss, usage = m(torch.zeros(0, 64, 8, 8, device=dev))
        torch.cuda.synchronize()
        assert q.shape == (0, 64, 8, 8) and idx.shape == (0, 8, 8) and idx.dtype == torch.int64
        assert loss.item() == 0.0 and usage.item() == 100.0
