"""Device times of the phases of the bench workload (100x60 grid, 24x16 rays, n LOS) on one GPU: development aid.
python tools/phase_times.py [n_los] [f64|f32] [reps]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")


def main():
    n_los = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    scn = synth.make_scenario(100, 60, 24, 16, n_em=1, rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5)
    G = binding.GpuModel(scn, prec)
    locs, dirs = synth.random_los(n_los)
    G.ctx.los_upload(G.ctx.los_from_MSO(locs, dirs))
    for it in range(reps):
        G.ctx.influence(0, scn.n_vox)
        tr, ma = G.ctx.kernel_ms(binding.PH_TRAVERSE)[0], G.ctx.kernel_ms(binding.PH_INFLUENCE)[0]
        G.ctx.solve()
        so = G.ctx.kernel_ms(binding.PH_SOLVE)[0]
        G.ctx.brightness_resident(10)
        lt, od, br = (G.ctx.kernel_ms(p)[0] for p in (binding.PH_TRAVERSE, binding.PH_ORDER, binding.PH_BRIGHTNESS))
        print(f"{prec} rep {it}: voxel-ray traverse {tr:.3f}  march {ma:.3f}  solve {so:.3f}  LOS traverse {lt:.3f}  "
              f"order {od:.3f}  brightness {br:.3f}  sum {tr + ma + so + lt + od + br:.3f} ms  residual {G.ctx.residual(0):.2e}",
              flush=True)


if __name__ == "__main__":
    main()
