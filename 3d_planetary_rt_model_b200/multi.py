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


def partition_interleaved(n: int, world: int, rank: int, chunks_per_rank: int = 8):
    """Source voxels for this rank as `chunks_per_rank` ranges dealt round-robin: [(lo, hi), ...], ascending.
    Rows are not equally expensive -- a voxel-origin ray from a low-altitude voxel crosses more cells than one from
    the outer corona (measured on the 100x60 grid: the lower half of the voxels costs 15 % more than the upper
    half) -- so contiguous halves leave one rank waiting; dealing chunks of n / (world * chunks_per_rank) voxels
    evens that out to ~1 %.  world = 1 gives the single range [(0, n)]."""
    if world == 1:
        return [(0, n)]
    total = world * chunks_per_rank
    out = []
    for c in range(rank, total, world):
        lo, hi = partition(n, total, c)
        if hi > lo:
            out.append((lo, hi))
    return out


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


def connect_row_sink(dist, ctx, rank: int, root: int = 0, n_emissions: int = 1):
    """Peer-memory form of the same exchange (include/b200rt.h, "multi-GPU row exchange over peer memory"): the
    solving rank exports a CUDA IPC handle of its resident K, every other rank opens it and names it as the sink of
    its row batches.  After this, ctx.influence(v0, v1) on a non-root rank DMAs its finished row batches into the
    root's K over NVLink while it marches the next batch; a dist.barrier() after the call releases the solve.
    Returns the peer pointers opened on this rank (close them with ctx.ipc_close before the root destroys its context).
    The handles travel through broadcast_object_list (host side, once per grid size)."""
    handles = [ctx.ipc_export_influence(e) for e in range(n_emissions)] if rank == root else [None] * n_emissions
    dist.broadcast_object_list(handles, src=root)
    ptrs = []
    if rank != root:
        for e, h in enumerate(handles):
            p = ctx.ipc_open(h)
            ctx.set_row_sink(e, p)
            ptrs.append(p)
    return ptrs


def connect_exchange(dist, ctx, rank: int, world: int):
    """Distributed solve (include/b200rt.h, "distributed solve"): every rank exports a CUDA IPC handle of its exchange
    block, every rank opens every other one.  Returns blocks[q] = rank q's block as addressable from this process
    (blocks[rank] is the own device pointer), the argument of ctx.solve_distributed(rank, world, blocks).  After this the
    ranks never meet on the host again: rows stay where they were built, S ends up resident everywhere.  Close the
    foreign entries with ctx.ipc_close before the contexts are destroyed."""
    ptr, handle = ctx.solve_exchange(want_ipc=world > 1)
    if world == 1:
        return [ptr]
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    return [ptr if q == rank else ctx.ipc_open(handles[q]) for q in range(world)]


def broadcast_vector(dist, v, root: int = 0):
    """the solved source function (n_vox doubles) from the solving rank to every rank"""
    dist.broadcast(v, src=root)
