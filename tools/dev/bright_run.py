"""Development aid: device time of the brightness phase on the bench workload (python tools/dev/bright_run.py [n_los])"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scn = synth.make_scenario(100, 60, 24, 16, n_em=1, rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5)
G = binding.GpuModel(scn, "f64")
G.ctx.set_sourcefn(0, np.linspace(1.0, 0.01, scn.n_vox))
locs, dirs = synth.random_los(n)
los = G.ctx.los_from_MSO(locs, dirs)
G.ctx.los_upload(los)
for i in range(3):
    G.ctx.brightness_resident(10)
    print("brightness ms", G.ctx.kernel_ms(binding.PH_BRIGHTNESS), "traverse ms", G.ctx.kernel_ms(binding.PH_TRAVERSE))
out = G.ctx.los_download()
print("checksum", float(np.sum(out["brightness"])))
