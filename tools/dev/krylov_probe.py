"""How many GMRES iterations does (I - w K) S = S0 take?  K from the CPU oracle (development probe, CPU only).
python tools/dev/krylov_probe.py n_rb n_sb n_theta n_phi [nH_exo] [out.npz]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
import oraclebind

def gmres(A, b, tol, m):
    """unrestarted GMRES with CGS2; returns x, list of relative residual estimates"""
    n = len(b)
    V = np.zeros((m + 1, n)); H = np.zeros((m + 1, m))
    beta = np.linalg.norm(b); V[0] = b / beta
    g = np.zeros(m + 1); g[0] = beta
    cs = np.zeros(m); sn = np.zeros(m); hist = []
    for j in range(m):
        w = A @ V[j]
        h = V[:j + 1] @ w; w -= V[:j + 1].T @ h
        h2 = V[:j + 1] @ w; w -= V[:j + 1].T @ h2; h += h2
        H[:j + 1, j] = h; H[j + 1, j] = np.linalg.norm(w); V[j + 1] = w / H[j + 1, j]
        for i in range(j):
            t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]; H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]; H[i, j] = t
        d = np.hypot(H[j, j], H[j + 1, j]); cs[j] = H[j, j] / d; sn[j] = H[j + 1, j] / d
        H[j, j] = d; H[j + 1, j] = 0
        g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]
        hist.append(abs(g[j + 1]) / beta)
        if hist[-1] < tol: break
    k = j + 1
    y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
    return V[:k].T @ y, hist

a = [int(x) for x in sys.argv[1:5]]
nH = float(sys.argv[5]) if len(sys.argv) > 5 else 5e5
kw = dict(rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5) if a[0] >= 100 else {}
scn = synth.make_scenario(*a, n_em=1, nH_exo=nH, **kw)
O = oraclebind.OracleModel(scn, "f64")
t0 = time.time(); O.build_rows(); print("K built in %.1f s, n = %d" % (time.time() - t0, scn.n_vox), flush=True)
K = O.K(0); S0 = O.vectors(0)["S0"]; w = float(scn.em_scalars[0][0])
if len(sys.argv) > 6: np.savez_compressed(sys.argv[6], K=K, S0=S0, w=w)
A = np.eye(scn.n_vox) - w * K
print("max row sum of wK %.4f, min %.4f; spectral radius (power it.) " % ((w * K).sum(1).max(), (w * K).sum(1).min()), end="")
v = np.ones(scn.n_vox)
for _ in range(200): v = K @ v; lam = np.linalg.norm(v); v /= lam
print("%.4f" % (w * lam))
x_lu = np.linalg.solve(A, S0)
for tol in (1e-8, 1e-10, 1e-12, 1e-14):
    x, hist = gmres(A, S0, tol, 120)
    print("tol %.0e: %d iterations, max rel err vs LU %.2e, true rel residual %.2e" % (tol, len(hist), np.max(np.abs(x - x_lu) / np.abs(x_lu)), np.linalg.norm(A @ x - S0) / np.linalg.norm(S0)))
print("residual history", " ".join("%.1e" % h for h in hist[:60]))
# Neumann / Jacobi for comparison
x = S0.copy()
for it in range(1, 2001):
    xn = S0 + w * (K @ x)
    d = np.max(np.abs(xn - x) / np.abs(xn)); x = xn
    if d < 1e-12: break
print("Neumann iterations to 1e-12 change: %d, err vs LU %.2e" % (it, np.max(np.abs(x - x_lu) / np.abs(x_lu))))
