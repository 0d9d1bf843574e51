"""Development aid: timeline of one solve on the bench grid (B200RT_SOLVE_TRACE=1 python tools/solve_trace.py)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
scn = synth.make_scenario(100, 60, 24, 16, n_em=1, rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5)
G = binding.GpuModel(scn, "f64")
G.build_rows()
for i in range(3):
    t0 = time.perf_counter(); r = G.solve(); t1 = time.perf_counter()
    print("solve wall ms", (t1 - t0) * 1e3, "kernel ms", G.ctx.kernel_ms(binding.PH_SOLVE), "residual", r, file=sys.stderr)
