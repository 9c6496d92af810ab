"""Seeded synthetic inputs for the golden vectors (shared by make_golden.py and the tests).

Every case is regenerated from its seed with the CPU generator; the golden file stores a
SHA-256 of the regenerated inputs so a drifted RNG is detected instead of silently compared.
Shapes follow SURVEY.md §8: C2 = (8,256,64,64) K=512; C1 = resnet50 levels at 2x3x512x512;
448-input grids (56^2, 28^2, 14^2); adversarial tie cases.
"""
import hashlib

import torch


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def _randn_case(seed, b, c, h, w, k):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, c, h, w, generator=g)
    e = torch.randn(k, c, generator=g)
    return x, e


def _relu_case(seed, b, c, h, w, k):
    """Realistic variant: non-negative features, codes = perturbed samples of x."""
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(b, c, h, w, generator=g))
    flat = x.permute(0, 2, 3, 1).reshape(-1, c)
    rows = torch.randperm(flat.shape[0], generator=g)[:k] if flat.shape[0] >= k else \
        torch.randint(0, flat.shape[0], (k,), generator=g)
    e = flat[rows] + 0.05 * torch.randn(k, c, generator=g)
    return x, e.contiguous()


def _uniform_init_case(seed, b, c, h, w, k):
    """kmeans_init=False initialisation: uniform(-1/K, 1/K) codebook (vq_img.py:157)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, c, h, w, generator=g)
    e = (torch.rand(k, c, generator=g) * 2 - 1) / k
    return x, e


def _adversarial(kind):
    g = torch.Generator().manual_seed(1234)
    if kind == "dup_codes":          # duplicated codebook rows -> lower index must win
        x, e = _randn_case(7, 2, 64, 8, 8, 96)
        e[50:60] = e[10:20]
        e[95] = e[0]
        return x, e
    if kind == "x_equals_code":      # x exactly equal to codes (distance 0 after clamp)
        x, e = _randn_case(8, 1, 64, 16, 16, 128)
        flat = x.permute(0, 2, 3, 1).reshape(-1, 64)
        e[:64] = flat[:64]
        e[100] = flat[3]             # duplicate of an exact hit at a higher index
        return x, e
    if kind == "equidistant":        # two codes equidistant by construction (+/- delta on one axis)
        c, k = 64, 64
        e = torch.randn(k, c, generator=g) * 4
        x = torch.zeros(1, c, 8, 8)
        base = torch.randn(64, c, generator=g) * 0.25
        d = torch.zeros(c); d[5] = 1.0
        e[7] = base[0] + d
        e[3] = base[0] - d
        e[40] = base[1] + 2 * d
        e[41] = base[1] - 2 * d
        flat = base.clone()
        x = flat.t().reshape(1, c, 8, 8).contiguous()
        return x, e
    if kind == "all_zero_x":
        x, e = _randn_case(9, 1, 64, 8, 8, 64)
        return torch.zeros_like(x), e
    if kind == "n_lt_k":             # fewer vectors than codes (N=16 < 25 and K=40 > 25 -> GEMM path)
        return _randn_case(10, 1, 32, 4, 4, 40)
    if kind == "tiny_exact_path":    # N<=25 and K<=25 -> ATen exact-difference path
        return _randn_case(11, 1, 16, 4, 4, 20)
    if kind == "k_not_tile":
        return _randn_case(12, 2, 128, 12, 12, 300)
    if kind == "k_small_odd":
        return _randn_case(13, 1, 48, 10, 10, 77)
    raise KeyError(kind)


# name -> (builder, modes)   modes: which outputs to record
FORWARD_CASES = {
    # headline microbench shape (BASELINE.json configs[1])
    "c2_randn":        (lambda: _randn_case(0, 8, 256, 64, 64, 512)),
    "c2_relu":         (lambda: _relu_case(1, 8, 256, 64, 64, 512)),
    # config 1: resnet50 levels 3..5 at 2x3x512x512
    "c1_l3":           (lambda: _randn_case(42, 2, 512, 64, 64, 512)),
    "c1_l4":           (lambda: _randn_case(43, 2, 1024, 32, 32, 512)),
    "c1_l5":           (lambda: _randn_case(44, 2, 2048, 16, 16, 512)),
    "c1_l3_relu":      (lambda: _relu_case(45, 2, 512, 64, 64, 512)),
    "c1_l5_uniform":   (lambda: _uniform_init_case(46, 2, 2048, 16, 16, 512)),
    # shipped training resolution 448 -> 56^2 / 28^2 / 14^2, batch 4 (not multiples of 128)
    "r448_l3":         (lambda: _relu_case(50, 4, 512, 56, 56, 512)),
    "r448_l4":         (lambda: _relu_case(51, 4, 1024, 28, 28, 512)),
    "r448_l5":         (lambda: _relu_case(52, 4, 2048, 14, 14, 512)),
    "odd_7x7":         (lambda: _randn_case(53, 3, 64, 7, 7, 32)),
    "d64":             (lambda: _randn_case(54, 2, 64, 24, 24, 256)),
    "d96_k1024":       (lambda: _randn_case(55, 1, 96, 32, 32, 1024)),
    "dup_codes":       (lambda: _adversarial("dup_codes")),
    "x_equals_code":   (lambda: _adversarial("x_equals_code")),
    "equidistant":     (lambda: _adversarial("equidistant")),
    "all_zero_x":      (lambda: _adversarial("all_zero_x")),
    "n_lt_k":          (lambda: _adversarial("n_lt_k")),
    "tiny_exact_path": (lambda: _adversarial("tiny_exact_path")),
    "k_not_tile":      (lambda: _adversarial("k_not_tile")),
    "k_small_odd":     (lambda: _adversarial("k_small_odd")),
}


def kmeans_case(name):
    """(samples_bchw, K, iters, init_indices, use_cosine)"""
    if name == "km_small":
        g = torch.Generator().manual_seed(60)
        x = torch.relu(torch.randn(2, 64, 32, 32, generator=g))
        idx = torch.randperm(2 * 32 * 32, generator=g)[:32]
        return x, 32, 10, idx, False
    if name == "km_c1_l5":
        g = torch.Generator().manual_seed(61)
        x = torch.relu(torch.randn(2, 1024, 16, 16, generator=g))
        idx = torch.randperm(512, generator=g)[:256]
        return x, 256, 10, idx, False
    if name == "km_one_iter":
        g = torch.Generator().manual_seed(62)
        x = torch.randn(2, 256, 32, 32, generator=g)
        idx = torch.randperm(2048, generator=g)[:128]
        return x, 128, 1, idx, False
    if name == "km_n_lt_k":          # N < K -> randint sampling with replacement (duplicates, empty clusters)
        g = torch.Generator().manual_seed(63)
        x = torch.randn(1, 32, 6, 6, generator=g)
        idx = torch.randint(0, 36, (48,), generator=g)
        return x, 48, 4, idx, False
    if name == "km_cosine":
        g = torch.Generator().manual_seed(64)
        x = torch.randn(2, 64, 16, 16, generator=g)
        idx = torch.randperm(512, generator=g)[:24]
        return x, 24, 6, idx, True
    raise KeyError(name)


KMEANS_CASES = ["km_small", "km_c1_l5", "km_one_iter", "km_n_lt_k", "km_cosine"]
COSINE_CASES = {"cos_small": (lambda: _randn_case(70, 2, 64, 16, 16, 48)),
                "cos_c2ish": (lambda: _relu_case(71, 2, 256, 32, 32, 512))}


# Round-2 cosine goldens (golden_cosine_v2.pt, make_golden_cosine.py): bit-exact indices in train AND eval mode, the
# normalised input and the renormalised weights pinned by SHA-256.  "cos2_dup" has duplicated codes (argmax ties ->
# first index) and a zero pixel (norm clamped at 1e-12); "cos2_d512" / "cos2_d1024" cross MKL's K-blocking thresholds
# (two half chains, 384-term blocks); "cos2_d100" has a channel count that is not a multiple of 8 (the tail rules of
# ATen's last-dim norm); "cos2_c2" is BASELINE config 2's shape.
def _cos_dup_case():
    x, e = _relu_case(72, 2, 96, 24, 24, 200)
    e[150:170] = e[20:40]
    e[199] = e[0]
    x[0, :, 3, 5] = 0.0
    return x, e.contiguous()


L2NORM_LASTDIM_DIMS = [1, 3, 4, 7, 8, 9, 12, 13, 15, 16, 23, 31, 64, 100, 103, 255, 511, 1000]
L2NORM_VIEW_SHAPES = [(1, 5, 7), (2, 33, 45), (3, 100, 131), (2, 300, 64), (1, 17, 1000)]

COSINE2_CASES = {"cos2_small": (lambda: _randn_case(73, 2, 64, 16, 16, 48)),
                 "cos2_relu": (lambda: _relu_case(74, 2, 256, 32, 32, 512)),
                 "cos2_dup": _cos_dup_case,
                 "cos2_d100": (lambda: _randn_case(75, 3, 100, 20, 12, 130)),
                 "cos2_d512": (lambda: _relu_case(76, 2, 512, 32, 32, 300)),
                 "cos2_d1024": (lambda: _relu_case(77, 2, 1024, 32, 32, 512)),
                 "cos2_b1": (lambda: _randn_case(78, 1, 128, 40, 40, 256)),
                 "cos2_c2": (lambda: _relu_case(79, 8, 256, 64, 64, 512))}


# VQ segmentation head: (seed, B, C=dim, H, W, K=classes, distance).  Decoder features are
# post-ReLU (non-negative); prototypes are perturbed feature rows.  "sh_dup" has a pixel equal to a prototype
# (distance ~0: catastrophic cancellation in the augmented form) and two identical prototypes (tie -> lower index).
# "sh_odd" is the single-image case (B = 1): there ATen skips cdist's contiguous copy and sums |x|^2 over the
# strided dim in its "outer" order (which depends on the host thread partition); distances then differ from the
# B >= 2 arithmetic by an ulp or two -- see DESIGN.md "known divergences".
def _seghead_case(seed, b, c, h, w, k, exact_hit=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(b, c, h, w, generator=g)) + 0.01
    flat = x.permute(0, 2, 3, 1).reshape(-1, c)
    e = flat[torch.randperm(flat.shape[0], generator=g)[:k]] + 0.3 * torch.randn(k, c, generator=g)
    if exact_hit:
        e[1] = flat[7]
        e[k - 1] = e[0]
    return x, e.contiguous()


SEGHEAD_CASES = {
    "sh_c3_d32": (lambda: _seghead_case(80, 2, 32, 32, 32, 3), "euclidean"),
    "sh_odd": (lambda: _seghead_case(81, 1, 16, 24, 40, 5), "euclidean"),
    "sh_k19_d64": (lambda: _seghead_case(82, 2, 64, 16, 16, 19), "euclidean"),
    "sh_dup": (lambda: _seghead_case(83, 2, 24, 8, 12, 4, exact_hit=True), "euclidean"),
    "sh_cos": (lambda: _seghead_case(84, 2, 32, 16, 16, 3), "cosine"),
}
