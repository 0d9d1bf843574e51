"""The peer-memory row exchange (include/b200rt.h "multi-GPU row exchange over peer memory", multi.connect_row_sink):
two PROCESSES, rank 1 marches the second half of the source voxels and its finished row batches are DMA'd into rank 0's
resident K through a CUDA IPC mapping; rank 0 marches the first half, then solves.  Result = the one-process run.
CUDA IPC needs two processes but not two devices, so this runs on the one-GPU test box (both ranks on cuda:0); on a
multi-GPU box rank 1 takes cuda:1 and the copies cross NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_two_process_row_push_equals_one_process(synth, binding, tmp_path):
    import ctypes
    ndev = binding.load().b200rt_device_count()
    devs = [0, 1 if ndev > 1 else 0]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "ipc_worker.py"), str(r), str(tmp_path), str(devs[r])],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    S = np.load(tmp_path / "S.npy")
    K0 = np.load(tmp_path / "K0.npy")
    scn = synth.make_scenario(40, 20, 7, 12, n_em=2, sza_T_contrast=0.1)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    G.solve()
    assert rel_err(G.K(0), K0, floor=1e-300) < 1e-12          # every row arrived (atomics: equal to rounding)
    for e in range(2):
        assert rel_err(G.vectors(e)["S"], S[e]) < 1e-10


@pytest.mark.gpu
def test_two_process_distributed_solve(synth, binding, tmp_path):
    """b200rt_solve_distributed between two PROCESSES (exchange blocks opened through CUDA IPC, as multi.connect_exchange
    does): each rank keeps its own rows, both end with the same S -- the one the LU of the full matrix gives"""
    ndev = binding.load().b200rt_device_count()
    devs = [0, 1 if ndev > 1 else 0]
    env = dict(os.environ)
    if ndev < 2:
        env["B200RT_KRYLOV_CTAS"] = "96"                     # both resident grids on one GPU
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "ipc_solve_worker.py"), str(r), str(tmp_path), str(devs[r])],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    S0, S1 = np.load(tmp_path / "S0.npy"), np.load(tmp_path / "S1.npy")
    m0, m1 = np.load(tmp_path / "meta0.npy"), np.load(tmp_path / "meta1.npy")
    assert np.array_equal(S0, S1) and np.array_equal(m0, m1)          # the ranks ran the same arithmetic
    assert 5 < m0[0] < 120 and max(m0[1:]) < 1e-12
    scn = synth.make_scenario(40, 20, 7, 12, n_em=2, sza_T_contrast=0.1)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    G.solve()
    for e in range(2):
        assert rel_err(G.vectors(e)["S"], S0[e], floor=1e-30) < 1e-7
