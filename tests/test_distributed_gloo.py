"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in vq_seg_b200.distributed: the collectives
and key packing are exercised for real; the per-rank kernels are replaced by the CPU oracle through the
`*_fn` hooks (tests may use the oracle; the product path never does)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_keys(x, cb, base):
    from vq_seg_b200.distributed import pack_keys
    d = O.euclidean_dist(x, cb)
    idx = torch.argmin(d, dim=-1)
    best = d.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    return pack_keys(best, idx + base)


def _unpack(keys, k_total):
    from vq_seg_b200.distributed import unpack_keys
    idx, d = unpack_keys(keys)
    return idx, d, torch.bincount(idx.reshape(-1), minlength=k_total)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from vq_seg_b200 import distributed as D
    g = torch.Generator().manual_seed(123)
    x = torch.randn(2, 600, 48, generator=g)               # (B, P, D), identical on both ranks
    e = torch.randn(96, 48, generator=g)
    e[70] = e[5]                                           # duplicate across shards: lower global index must win
    out = {}
    # --- codebook-sharded assignment: K split 48 / 48
    kl = 96 // world
    idx, dd, counts = D.sharded_assign(x, e[rank * kl:(rank + 1) * kl], rank * kl, 96,
                                       local_keys_fn=_oracle_keys, unpack_fn=_unpack)
    out["sharded"] = (idx, dd, counts)
    # --- data-parallel stats: rows split by rank, ONE exchange step
    flat = x.reshape(-1, 48)
    n = flat.shape[0]
    b, en = D.shard_rows(n)
    means0 = flat[:32].clone()
    buckets = torch.argmin(O.euclidean_dist(flat[b:en].unsqueeze(0), means0)[0], dim=-1)
    cnt = torch.bincount(buckets, minlength=32)
    sums = torch.zeros(32, 48).scatter_add_(0, buckets.unsqueeze(1).expand(-1, 48).contiguous(), flat[b:en])
    D.allreduce_code_stats(cnt, sums)
    out["stats"] = (cnt, sums)
    # --- data-parallel EMA extension: the all-reduced statistics feed the same update on every rank
    out["ema"] = O.ema_update(cnt, sums, torch.full((32,), 3.0), means0.clone(), 0.8, 1e-5)
    # --- identical k-means start on every rank from rank 0's global choice
    ids = torch.randperm(n, generator=torch.Generator().manual_seed(7 + rank))[:16]      # ranks DISAGREE on purpose
    rows = D.dp_init_means(flat[b:en].unsqueeze(0), b, ids,
                           gather_rows_fn=lambda xx, ii: torch.where((ii >= 0).unsqueeze(1), xx[0][ii.clamp(min=0)],
                                                                     torch.zeros(1)))
    out["init"] = (ids, rows)
    out["usage"] = D.global_code_usage(torch.bincount(buckets, minlength=32),
                                       usage_fn=lambda c: O.code_usage_from_counts(c, 32))
    results[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_modes():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    r0, r1 = results[0], results[1]
    g = torch.Generator().manual_seed(123)
    x = torch.randn(2, 600, 48, generator=g)
    e = torch.randn(96, 48, generator=g)
    e[70] = e[5]
    # sharded == single-device oracle, bit-exact indices / counts, same on both ranks
    ref_idx = O.assign_euclidean(x, e)
    for r in (r0, r1):
        idx, dd, counts = r["sharded"]
        assert torch.equal(idx, ref_idx)
        assert torch.equal(counts, torch.bincount(ref_idx.reshape(-1), minlength=96))
        assert not (idx == 70).any()                        # tie across shards resolved to the lower index
    # data-parallel stats == single-device stats
    flat = x.reshape(-1, 48)
    buckets = torch.argmin(O.euclidean_dist(flat.unsqueeze(0), flat[:32])[0], dim=-1)
    cnt = torch.bincount(buckets, minlength=32)
    sums = torch.zeros(32, 48).scatter_add_(0, buckets.unsqueeze(1).expand(-1, 48).contiguous(), flat)
    for r in (r0, r1):
        assert torch.equal(r["stats"][0], cnt)               # counts bit-exact for any rank count
        torch.testing.assert_close(r["stats"][1], sums, rtol=1e-5, atol=1e-5)
    # EMA extension: both ranks end with the codebook a single process computes from all rows
    cs_ref, ea_ref, w_ref = O.ema_update(cnt, sums, torch.full((32,), 3.0), flat[:32].clone(), 0.8, 1e-5)
    for r in (r0, r1):
        assert torch.equal(r["ema"][0], cs_ref)
        torch.testing.assert_close(r["ema"][2], w_ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(r0["ema"][2], r1["ema"][2])
    # identical start, taken from rank 0's ids
    ids0 = torch.randperm(1200, generator=torch.Generator().manual_seed(7))[:16]
    assert torch.equal(r0["init"][1], flat[ids0]) and torch.equal(r1["init"][1], flat[ids0])
    torch.testing.assert_close(r0["usage"], O.code_usage_from_counts(cnt, 32))
    assert torch.equal(r0["usage"], r1["usage"])


def test_key_packing_orders_like_argmin():
    from vq_seg_b200.distributed import pack_keys, unpack_keys
    d = torch.tensor([0.0, 1.5, 1.5, 3e-39, 2.0 ** 100, float("inf")])
    i = torch.tensor([7, 3, 9, 1, 0, 4])
    k = pack_keys(d, i)
    assert (k >= 0).all()
    order = torch.argsort(k)
    assert order.tolist() == [0, 3, 1, 2, 4, 5]               # by distance, ties by lower index
    ii, dd = unpack_keys(k)
    assert torch.equal(ii, i) and torch.equal(dd, d)
