import importlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
scn = synth.make_scenario(100, 60, 24, 16, n_em=1, rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5)
G = binding.GpuModel(scn, "f64")
G.ctx.influence(0, scn.n_vox)
G.ctx.solve(); G.ctx.solve()
print("graph solve ms", G.ctx.kernel_ms(binding.PH_SOLVE))
os.environ["B200RT_SOLVE_TRACE"] = "1"
G.ctx.solve()
print("traced (no graph) solve ms", G.ctx.kernel_ms(binding.PH_SOLVE))
