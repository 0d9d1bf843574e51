"""One rank of the two-process distributed-solve test (tests/test_multi_ipc.py): python ipc_solve_worker.py rank dir [device]
Each rank builds only its interleaved shard of the influence rows and keeps it; the exchange blocks are opened through
CUDA IPC (b200rt_solve_exchange -> b200rt_ipc_open), exactly what multi.connect_exchange does over torch.distributed.
Rendezvous through files in `dir`.  Both ranks may sit on the same GPU (then their resident grids must both fit)."""
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


def publish(path, data=b""):
    with open(path + ".tmp", "wb") as f:
        f.write(data)
    os.rename(path + ".tmp", path)


def main():
    rank, d = int(sys.argv[1]), sys.argv[2]
    device = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    multi = importlib.import_module(PKG + ".multi")
    scn = synth.make_scenario(40, 20, 7, 12, n_em=2, sza_T_contrast=0.1)
    G = binding.GpuModel(scn, "f64", device=device)
    own, handle = G.ctx.solve_exchange(want_ipc=True)
    publish(os.path.join(d, f"handle{rank}"), handle)
    wait_for(os.path.join(d, f"handle{1 - rank}"))
    other = G.ctx.ipc_open(open(os.path.join(d, f"handle{1 - rank}"), "rb").read())
    blocks = [own, other] if rank == 0 else [other, own]
    G.ctx.influence_ranges(multi.partition_interleaved(scn.n_vox, 2, rank, 4))
    for _ in range(2):                                      # twice: the round counters carry on
        G.ctx.solve_distributed(rank, 2, blocks)
    np.save(os.path.join(d, f"S{rank}.npy"), np.stack([G.vectors(e)["S"] for e in range(2)]))
    np.save(os.path.join(d, f"meta{rank}.npy"), np.array([G.ctx.last_solve_steps(), G.ctx.residual(0), G.ctx.residual(1)]))
    publish(os.path.join(d, f"done{rank}"))
    wait_for(os.path.join(d, f"done{1 - rank}"))            # the peer's block stays mapped until it has finished
    G.ctx.ipc_close(other)


if __name__ == "__main__":
    main()
