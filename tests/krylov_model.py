"""The distributed solve of csrc/solve_krylov.cu restated in numpy over torch.distributed (TEST INFRASTRUCTURE: a model
of the algorithm and of what the ranks exchange, used by the CPU suite under gloo; the product path is the CUDA one).

Every rank holds only ITS rows of K (interleaved shards, multi.partition_interleaved).  Exchanged: one all-gather of the
right-hand side rows, one of the preconditioner's block entries, and per GMRES step one all-gather of the pieces of the
product -- nothing else; the orthogonalisation (classical Gram-Schmidt twice, |w|^2 = |w1|^2 - sum h2^2), the Givens
rotations and the stopping test run redundantly on every rank with identical results.  Right preconditioner: the inverted
diagonal blocks of A = I - w K over the SZA columns (voxel = i_r * n_col + i_col)."""
import numpy as np
import torch


def _assemble(dist, world, rows, values, n):
    """all ranks contribute values for their rows -> the full vector (or [n, m] array) on every rank"""
    values = np.asarray(values, dtype=np.float64)
    tail = values.shape[1:]
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(rows)], dtype=torch.int64))
    m = int(max(c.item() for c in counts))
    pad_r = np.full(m, -1, dtype=np.int64); pad_r[:len(rows)] = rows
    pad_v = np.zeros((m,) + tail); pad_v[:len(rows)] = values
    got_r = [torch.zeros(m, dtype=torch.int64) for _ in range(world)]
    got_v = [torch.zeros((m,) + tail, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(got_r, torch.from_numpy(pad_r))
    dist.all_gather(got_v, torch.from_numpy(pad_v))
    out = np.zeros((n,) + tail)
    census = 0
    for r, v in zip(got_r, got_v):
        r, v = r.numpy(), v.numpy()
        keep = r >= 0
        out[r[keep]] = v[keep]
        census += int(keep.sum())
    if census != n:
        raise RuntimeError("the rows the ranks built do not add up to the grid")
    return out


def solve(dist, world, rows, K_rows, branching, S0, n_r, n_col, tol=1e-13, max_it=160, precondition=True):
    """rows: the voxels this rank built; K_rows[k] = K[rows[k], :].  -> (S, steps, true relative residual)"""
    n = n_r * n_col
    rows = np.asarray(rows, dtype=np.int64)
    A_rows = -branching * K_rows
    A_rows[np.arange(len(rows)), rows] += 1.0
    b = _assemble(dist, world, rows, S0[rows], n)
    Minv = None
    B_rows = A_rows
    if precondition:
        # block entries of the own rows: A[v][(i', column of v)]
        cols_of = lambda v: np.arange(n_r) * n_col + (v % n_col)
        entries = np.stack([A_rows[k, cols_of(v)] for k, v in enumerate(rows)]) if len(rows) else np.zeros((0, n_r))
        pre = _assemble(dist, world, rows, entries, n)                       # [n][n_r] on every rank
        Minv = [np.linalg.inv(pre[np.arange(n_r) * n_col + j]) for j in range(n_col)]
        B_rows = np.array(A_rows)
        for j in range(n_col):
            c = np.arange(n_r) * n_col + j
            B_rows[:, c] = A_rows[:, c] @ Minv[j]
    beta = np.sqrt(np.sum(b * b))
    if beta == 0:
        return np.zeros(n), 0, 0.0
    V = [b / beta]
    w_raw, inv_norm = b, 1.0 / beta
    g = np.zeros(max_it + 1); g[0] = beta
    cs, sn = np.zeros(max_it), np.zeros(max_it)
    R = np.zeros((max_it + 1, max_it))
    k = 0
    for j in range(max_it):
        x = w_raw * inv_norm
        w = _assemble(dist, world, rows, B_rows @ x, n)
        Vm = np.array(V)
        h1 = Vm @ w
        w = w - Vm.T @ h1
        w1sq = np.sum(w * w)
        h2 = Vm @ w
        w = w - Vm.T @ h2
        hn = np.sqrt(max(w1sq - np.sum(h2 * h2), 0.0))
        h = np.concatenate([h1 + h2, [0.0]])
        hi = h[0]
        for i in range(j):
            up = h[i + 1]
            h[i] = cs[i] * hi + sn[i] * up
            hi = -sn[i] * hi + cs[i] * up
        d = np.hypot(hi, hn)
        cs[j], sn[j] = (hi / d, hn / d) if d > 0 else (1.0, 0.0)
        h[j] = d
        R[:j + 1, j] = h[:j + 1]
        g[j + 1] = -sn[j] * g[j]
        g[j] = cs[j] * g[j]
        k = j + 1
        w_raw, inv_norm = w, (1.0 / hn if hn > 0 else 0.0)
        V.append(w * inv_norm)
        if abs(g[j + 1]) / beta <= tol or hn == 0:
            break
    y = np.linalg.solve(np.triu(R[:k, :k]), g[:k])
    u = np.array(V[:k]).T @ y
    S = u
    if precondition:
        S = np.zeros(n)
        for jcol in range(n_col):
            c = np.arange(n_r) * n_col + jcol
            S[c] = Minv[jcol] @ u[c]
    AS = _assemble(dist, world, rows, A_rows @ S, n)
    return S, k, float(np.sqrt(np.sum((b - AS) ** 2)) / beta)
