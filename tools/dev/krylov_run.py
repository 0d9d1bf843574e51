"""bench-size (100x60) solve on one GPU: the dense LU against the distributed GMRES with world = 1 (development aid)
python tools/dev/krylov_run.py [n_rb n_sb n_theta n_phi]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
a = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else [100, 60, 24, 16]
kw = dict(rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5) if a[0] >= 100 else {}
scn = synth.make_scenario(*a, n_em=1, **kw)
G = binding.GpuModel(scn, "f64")
G.build_rows()
os.environ["B200RT_SOLVER"] = "lu"            # (b200rt_solve would take the GMRES itself at this size)
for it in range(3):
    t0 = time.perf_counter(); G.ctx.solve(); w = time.perf_counter() - t0
    print("LU    : device %.3f ms, wall %.3f ms, residual %.2e" % (G.ctx.kernel_ms(binding.PH_SOLVE)[0], w * 1e3, G.ctx.residual(0)), flush=True)
S_lu = G.vectors(0)["S"].copy()
del os.environ["B200RT_SOLVER"]
block, _ = G.ctx.solve_exchange()
for it in range(3):
    t0 = time.perf_counter(); G.ctx.solve_distributed(0, 1, [block]); w = time.perf_counter() - t0
    ms, launches = G.ctx.kernel_ms(binding.PH_SOLVE)
    print("GMRES : device %.3f ms, wall %.3f ms, %d steps, %d launches, true residual %.2e" % (ms, w * 1e3, G.ctx.last_solve_steps(), launches, G.ctx.residual(0)), flush=True)
S = G.vectors(0)["S"]
print("max element-wise relative difference to the LU solution: %.2e" % np.max(np.abs(S - S_lu) / np.abs(S_lu)))
