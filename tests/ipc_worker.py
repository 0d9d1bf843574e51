"""One rank of the two-process peer-memory row exchange test (tests/test_multi_ipc.py): python ipc_worker.py rank dir [device]
Both ranks may sit on the SAME GPU: CUDA IPC works between processes whatever the device, so the exchange is testable
on a one-GPU box.  Rendezvous through files in `dir` (the bench uses torch.distributed for the same hand-over)."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_planetary_rt_model_b200"


def wait_for(path, timeout=120.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)


def main():
    rank, d = int(sys.argv[1]), sys.argv[2]
    device = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    multi = importlib.import_module(PKG + ".multi")
    scn = synth.make_scenario(40, 20, 7, 12, n_em=2, sza_T_contrast=0.1)
    G = binding.GpuModel(scn, "f64", device=device)
    n = scn.n_vox
    ranges = multi.partition_interleaved(n, 2, rank, 4)      # four interleaved shards per rank
    if rank == 0:
        hs = [G.ctx.ipc_export_influence(e) for e in range(2)]
        with open(os.path.join(d, "handles.tmp"), "wb") as f:
            f.write(b"".join(hs))
        os.rename(os.path.join(d, "handles.tmp"), os.path.join(d, "handles"))
        G.ctx.influence_ranges(ranges)
        wait_for(os.path.join(d, "rank1_done"))
        G.ctx.solve()
        np.save(os.path.join(d, "S.npy"), np.stack([G.vectors(e)["S"] for e in range(2)]))
        np.save(os.path.join(d, "K0.npy"), G.K(0))
        open(os.path.join(d, "rank0_done"), "w").close()
    else:
        wait_for(os.path.join(d, "handles"))
        raw = open(os.path.join(d, "handles"), "rb").read()
        ptrs = [G.ctx.ipc_open(raw[64 * e:64 * (e + 1)]) for e in range(2)]
        for e, p in enumerate(ptrs):
            G.ctx.set_row_sink(e, p)
        G.ctx.influence_ranges(ranges)       # returns when the rows have landed in rank 0's K
        open(os.path.join(d, "rank1_done"), "w").close()
        wait_for(os.path.join(d, "rank0_done"))
        for p in ptrs:
            G.ctx.ipc_close(p)


if __name__ == "__main__":
    main()
