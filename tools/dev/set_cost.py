"""Development aid: where one parameter set of the sweep spends its time (grid 40x20x7x12, 2 emissions, 1e4 LOS)"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
scn = synth.make_scenario()
G = binding.GpuModel(scn, "f64")
locs, dirs = synth.random_los(10000)
los = G.ctx.los_from_MSO(locs, dirs)
G.ctx.los_upload(los)
for it in range(4):
    t0 = time.perf_counter(); G.ctx.influence(0, scn.n_vox); t1 = time.perf_counter()
    ki = (G.ctx.kernel_ms(binding.PH_TRAVERSE), G.ctx.kernel_ms(binding.PH_INFLUENCE))
    G.ctx.solve(); t2 = time.perf_counter()
    ks = G.ctx.kernel_ms(binding.PH_SOLVE)
    G.ctx.brightness_resident(10); t3 = time.perf_counter()
    kb = (G.ctx.kernel_ms(binding.PH_TRAVERSE), G.ctx.kernel_ms(binding.PH_BRIGHTNESS))
    print(f"influence wall {1e3*(t1-t0):.3f} ms kernels trav {ki[0]} march {ki[1]}; solve wall {1e3*(t2-t1):.3f} kernels {ks}; "
          f"brightness wall {1e3*(t3-t2):.3f} kernels trav {kb[0]} march {kb[1]}")
