"""Multi-GPU parity (needs >= 2 visible GPUs): NCCL data-parallel k-means and codebook-sharded assignment against the
single-GPU result, at 1 M rows.  The check itself lives in scripts/multi_gpu_parity.py so that it also runs
  * under torch.distributed.run:  python -m torch.distributed.run --nproc-per-node 2 ... scripts/multi_gpu_parity.py
  * inside `bench.py --gpus N` (N > 1), whose JSON line carries the result as extras.parity
-- the driver's pytest lease has one GPU, its scaling lease has eight."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, rows, results):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from multi_gpu_parity import run_parity
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    results[rank] = run_parity(rank, world, dev, rows=rows)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rows", [1 << 20])
def test_two_gpu_modes(rows):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (bench.py --gpus N runs the same check in the scaling lease: extras.parity)")
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), rows, results), nprocs=world, join=True)
    for r in range(world):
        out = results[r]
        assert out["rows"] >= 1 << 20
        assert out["sharded_idx_equal"] and out["sharded_counts_equal"], "sharded (min,index) reduction differs from the single-GPU argmin"
        assert out["dp_kmeans_bins_equal"], "data-parallel k-means counts differ from single-GPU"
        assert out["dp_kmeans_means_rel_err"] < 1e-5, out
        assert out["module_kmeans_hook_same_codebook"], "kmeans_reduce_fn: the ranks' codebooks differ after the k-means init"
        assert out["single_gpu_matches_brute_force"] and out["all_ranks_ok"]
