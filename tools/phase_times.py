"""device time of each phase on the bench workload (grid 100x60x24x16): python tools/phase_times.py [n_los]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
synth = importlib.import_module(bench.PKG + ".synth")
binding = importlib.import_module(bench.PKG + ".binding")
n_los = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scn, locs, dirs = bench.make_workload(synth, n_los)
ctx = binding.Context(0, binding.F64)
ctx.set_grid(ctx.make_grid(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, scn.rb, scn.szamethod, scn.raymethod))
ctx.set_singlet(0, 1, *(float(x) for x in scn.em_scalars[0]), binding.define_singlet_tables(scn, 0))
ctx.los_upload(ctx.los_from_MSO(locs, dirs))
for it in range(4):
    ctx.influence()
    a = (ctx.kernel_ms(binding.PH_TRAVERSE)[0], ctx.kernel_ms(binding.PH_INFLUENCE)[0])
    ctx.solve()
    b = ctx.kernel_ms(binding.PH_SOLVE)[0]
    ctx.brightness_resident(10)
    c = (ctx.kernel_ms(binding.PH_TRAVERSE)[0], ctx.kernel_ms(binding.PH_BRIGHTNESS)[0])
print(f"influence: traverse {a[0]:.3f} ms  march {a[1]:.3f} ms | solve {b:.3f} ms | brightness: traverse {c[0]:.3f} ms  march {c[1]:.3f} ms"
      f" | steps {ctx.last_step_count()} substeps {ctx.last_substep_count()}")
