"""Multi-GPU parity of the two exchange steps of the VQ path, sized for real inputs:

  * data-parallel k-means (vq_img.py:29-63): rows sharded over the ranks, one SUM all-reduce of counts + sums per
    Lloyd iteration -- against the single-GPU k-means of the whole input on rank 0 (bins bit-exact, means 1e-5);
  * codebook-sharded assignment: every rank scores all rows against its slice of the codebook, one MIN all-reduce of
    (distance bits << 32 | global index) keys -- against the single-GPU assignment (indices and counts bit-exact).

Run it one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_parity.py [--rows 1048576]
bench.py calls run_parity() in its --gpus N > 1 runs (extras.parity) and tests/test_gpu_multi.py spawns it.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_parity(rank, world, dev, rows=1 << 20, dim=128, k_kmeans=256, k_sharded=4096, iters=3):
    """Returns a dict of booleans / errors (identical on every rank).  Needs an initialised process group."""
    import vq_seg_b200 as V
    from vq_seg_b200 import ops, distributed as D
    images = 8 * world                                       # images of rows / images pixels each, NCHW like the encoder's maps
    pix = rows // images
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.relu(torch.randn(images, dim, pix, generator=g, device=dev))        # same tensor on every rank (same seed)
    xv = x.permute(0, 2, 1)
    out = {"rows": images * pix, "dim": dim, "world": world}
    # ---- data-parallel k-means: images split across ranks
    per = images // world
    x_local = xv[rank * per:(rank + 1) * per]
    init = torch.randperm(images * pix, generator=torch.Generator().manual_seed(9))[:k_kmeans].to(dev)
    # one iteration: the assignment uses the (identical) start means, so the counts must agree bit for bit and the
    # means differ only by the association order of the fp32 all-reduce; several iterations: a row within ~1e-7 of
    # a cell boundary may change sides once the means differ in the last bits (and a moved row shifts its cluster's
    # mean by 1/count), so later iterations are compared with a tolerance of one row per thousand (measured on two
    # B200s at 1 M rows: 70 rows moved after 3 iterations)
    means, bins = D.dp_kmeans(x_local, k_kmeans, 1, init, rank * per * pix)
    m1, b1 = V.kmeans(xv, k_kmeans, 1, init_indices=init)
    out["dp_kmeans_bins_equal"] = bool(torch.equal(bins, b1))
    out["dp_kmeans_means_rel_err"] = float(((means - m1).abs().max() / m1.abs().max()).item())
    means, bins = D.dp_kmeans(x_local, k_kmeans, iters, init, rank * per * pix)
    m1, b1 = V.kmeans(xv, k_kmeans, iters, init_indices=init)
    out["dp_kmeans_bins_moved_after_%d_iters" % iters] = int((bins - b1).abs().sum().item())
    out["dp_kmeans_means_rel_err_after_%d_iters" % iters] = float(((means - m1).abs().max() / m1.abs().max()).item())
    # ---- codebook-sharded assignment: K split across ranks, rows replicated
    e = torch.randn(k_sharded, dim, generator=g, device=dev)
    kl = k_sharded // world
    idx, dd, counts = D.sharded_assign(xv, e[rank * kl:(rank + 1) * kl].contiguous(), rank * kl, k_sharded)
    ref_idx, ref_counts = ops.assign(xv, e, ops.prepare_codebook(e), ops.ALGO_AUTO)
    out["sharded_idx_equal"] = bool(torch.equal(idx, ref_idx))
    out["sharded_counts_equal"] = bool(torch.equal(counts, ref_counts))
    sub = xv[:1, :4096]                                       # the single-GPU assignment itself against brute force
    ex_idx, _ = ops.assign(sub, e, None, ops.ALGO_EXACT)
    out["single_gpu_matches_brute_force"] = bool(torch.equal(ex_idx, ref_idx[:1, :4096]))
    # ---- the module's own k-means hook (kmeans_init=True + codebook.kmeans_reduce_fn): every rank sees different
    # images and draws its own start rows; rank 0's start is broadcast and the statistics are all-reduced, so the
    # replicated codebooks must come out identical on all ranks after the first training forward
    torch.manual_seed(100 + rank)                                                  # ranks DISAGREE on the RNG on purpose
    m = V.VectorQuantizer(dim=dim, num_embeddings=64, kmeans_init=True, kmeans_iters=3).to(dev)
    m.codebook.kmeans_reduce_fn = D.allreduce_code_stats
    m.train()
    m(x[rank * per:(rank + 1) * per].reshape(per, dim, 64, pix // 64)[:, :, :, :64].contiguous())
    w = m.codebook.embedding.weight.detach()
    w0 = w.clone()
    dist.broadcast(w0, src=0)
    out["module_kmeans_hook_same_codebook"] = bool(torch.equal(w, w0))
    flags = torch.tensor([int(out["module_kmeans_hook_same_codebook"]), int(out["dp_kmeans_bins_equal"]), int(out["sharded_idx_equal"]), int(out["sharded_counts_equal"]),
                          int(out["single_gpu_matches_brute_force"]), int(out["dp_kmeans_means_rel_err"] < 1e-5),
                          int(out["dp_kmeans_bins_moved_after_%d_iters" % iters] <= max(2, out["rows"] // 1000))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(flags.min().item() == 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    out = run_parity(rank, world, dev, rows=args.rows)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
