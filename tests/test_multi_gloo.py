"""The N > 1 host logic on CPU: world size 2 (and 3) over gloo -- row / LOS partitioning and the one
exchange of the pipeline (row blocks gathered onto the solving rank), ragged and empty blocks included."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "3d_planetary_rt_model_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, n_cols, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    multi = importlib.import_module(PKG + ".multi")
    lo, hi = multi.partition(n_rows, world, rank)
    K = torch.zeros(n_rows, n_cols, dtype=torch.float64)
    full = torch.arange(n_rows * n_cols, dtype=torch.float64).reshape(n_rows, n_cols) + 1.0
    K[lo:hi] = full[lo:hi]                     # "this rank built rows [lo, hi)"
    multi.gather_rows(dist, K, n_rows, rank, world, 0)
    S = torch.zeros(n_rows, dtype=torch.float64)
    if rank == 0:
        assert torch.equal(K, full)
        S = K.sum(dim=1)                       # stand-in for the solve on the gathering rank
    multi.broadcast_vector(dist, S, 0)
    assert torch.equal(S, full.sum(dim=1))
    # lines of sight: disjoint slices whose union is everything, no communication
    l0, l1 = multi.partition(1000003, world, rank)
    counts = torch.tensor([l1 - l0], dtype=torch.int64)
    dist.all_reduce(counts)
    assert int(counts) == 1000003
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([lo, hi]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 77), (2, 1), (3, 10)])
def test_row_gather_and_broadcast(tmp_path, world, n_rows):
    mp.spawn(_worker, args=(world, _free_port(), n_rows, 13, str(tmp_path)), nprocs=world, join=True)
    blocks = [np.load(tmp_path / f"ok{r}.npy") for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n_rows
    for a, b in zip(blocks[:-1], blocks[1:]):
        assert a[1] == b[0]


def test_partition_properties():
    multi = importlib.import_module(PKG + ".multi")
    for n in (0, 1, 7, 741, 5841, 10**6):
        for world in (1, 2, 3, 4, 8):
            parts = [multi.partition(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and all(s >= 0 for s in sizes)
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
