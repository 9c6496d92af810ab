"""Multi-GPU parity (needs >= 2 visible GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
NCCL data-parallel k-means and codebook-sharded assignment vs the single-GPU result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import vq_seg_b200 as V
    from vq_seg_b200 import ops, distributed as D
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(4, 128, 1024, generator=g)).to(dev)           # (B, D, HW) NCHW-like
    e = torch.randn(512, 128, generator=g).to(dev)
    xv = x.permute(0, 2, 1)
    # sharded assignment: K split across ranks
    kl = 512 // world
    idx, dd, counts = D.sharded_assign(xv, e[rank * kl:(rank + 1) * kl].contiguous(), rank * kl, 512)
    ref_idx, ref_counts = ops.assign(xv, e, None, ops.ALGO_EXACT)
    ok_sharded = torch.equal(idx, ref_idx) and torch.equal(counts, ref_counts)
    # data-parallel k-means: images split across ranks
    per = 4 // world
    x_local = xv[rank * per:(rank + 1) * per]
    init = torch.randperm(4096, generator=torch.Generator().manual_seed(9))[:64].to(dev)
    means, bins = D.dp_kmeans(x_local, 64, 5, init, rank * per * 1024)
    m1, b1 = V.kmeans(xv, 64, 5, init_indices=init)
    ok_bins = torch.equal(bins, b1)
    err = ((means - m1).abs().max() / m1.abs().max()).item()
    results[rank] = (ok_sharded, ok_bins, err)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_modes():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    for r in range(world):
        ok_sharded, ok_bins, err = results[r]
        assert ok_sharded, "sharded (min,index) reduction differs from the single-GPU argmin"
        assert ok_bins, "data-parallel k-means counts differ from single-GPU"
        assert err < 1e-5, err
