"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Only what the path needs: how source-voxel rows and lines of sight are split over ranks, and the ONE
exchange of the pipeline -- every rank's block of influence-matrix rows gathered onto the solving
rank's resident K.  The functions work on any torch tensors, so the same code runs
under `nccl` on the B200s and under `gloo` in the CPU tests.  The reference has no multi-GPU code at
all (single cudaSetDevice(0), RT_gpu.cu:143,257)."""
from __future__ import annotations


def partition(n: int, world: int, rank: int):
    """contiguous block [lo, hi) of n items for this rank; sizes differ by at most one"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(dist, K, n_rows: int, rank: int, world: int, root: int = 0):
    """K: [n_rows, n_cols] tensor on every rank, rows partition(n_rows, world, r) valid on rank r.
    After the call rank `root` holds every row.  One grouped send/recv; ragged and empty blocks allowed."""
    ops = []
    if rank == root:
        for r in range(world):
            a, b = partition(n_rows, world, r)
            if r != root and b > a:
                ops.append(dist.P2POp(dist.irecv, K[a:b], r))
    else:
        a, b = partition(n_rows, world, rank)
        if b > a:
            ops.append(dist.P2POp(dist.isend, K[a:b], root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def broadcast_vector(dist, v, root: int = 0):
    """the solved source function (n_vox doubles) from the solving rank to every rank"""
    dist.broadcast(v, src=root)
