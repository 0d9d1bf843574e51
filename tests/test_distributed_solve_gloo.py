"""The distributed solve's algorithm and exchange pattern on CPU: world size 2 and 3 over gloo (tests/krylov_model.py, a
numpy restatement of csrc/solve_krylov.cu), with K from the oracle.  Each rank holds only its interleaved shard of the
rows; the result must be the dense solution, the same on every rank, in about half the steps with the column-block
preconditioner.  The CUDA form is tested by tests/test_distributed_solve.py / test_multi_ipc.py on the GPU box."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = "3d_planetary_rt_model_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, drop_rows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, HERE)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = importlib.import_module(PKG + ".synth")
    multi = importlib.import_module(PKG + ".multi")
    import krylov_model
    import oraclebind
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1, sza_T_contrast=0.1)
    O = oraclebind.OracleModel(scn, "f64")
    O.build_rows()
    K, S0, w = O.K(0), O.vectors(0)["S0"], float(scn.em_scalars[0][0])
    n, n_r, n_col = scn.n_vox, scn.n_rb - 1, scn.n_sb - 1
    rows = [v for a, b in multi.partition_interleaved(n, world, rank, 3) for v in range(a, b)]
    if drop_rows and rank == world - 1:
        rows = rows[:-2]
    res = {}
    try:
        for pc in (False, True):
            S, steps, resid = krylov_model.solve(dist, world, rows, K[rows], w, S0, n_r, n_col, precondition=pc)
            res[pc] = (S, steps, resid)
        exact = np.linalg.solve(np.eye(n) - w * K, S0)
        for pc in (False, True):
            assert np.max(np.abs(res[pc][0] - exact) / np.abs(exact)) < 1e-7, pc
            assert res[pc][2] < 1e-12
        assert res[True][1] < res[False][1]
        np.save(os.path.join(out_dir, f"S{rank}.npy"), res[True][0])
        np.save(os.path.join(out_dir, f"steps{rank}.npy"), np.array([res[False][1], res[True][1]]))
    except RuntimeError as ex:
        with open(os.path.join(out_dir, f"error{rank}"), "w") as f:
            f.write(str(ex))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_with_their_own_rows_reach_the_dense_solution(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), False), nprocs=world, join=True)
    S = [np.load(tmp_path / f"S{r}.npy") for r in range(world)]
    steps = [np.load(tmp_path / f"steps{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(S[0], S[r])                  # the ranks ran the same arithmetic on the same exchanged pieces
        assert np.array_equal(steps[0], steps[r])


def test_missing_rows_are_caught_by_the_census(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), True), nprocs=2, join=True)
    for r in range(2):
        assert "do not add up" in open(tmp_path / f"error{r}").read()
