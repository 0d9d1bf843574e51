import sys, importlib, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
from oracle import refbind
from util import rel_err
prec = "f32"
scn = synth.make_scenario(12, 8, 5, 6, n_em=2, sza_T_contrast=0.1)
cpu = refbind.RefModel(scn, prec, variant="b200"); gpu = refbind.RefModel(scn, prec, variant="b200")
cpu.generate_S(); gpu.generate_S_gpu()
G = binding.GpuModel(scn, prec); G.ctx.generate_S()
for e in range(2):
    a, b, c = cpu.vectors(e), gpu.vectors(e), G.vectors(e)
    for k in ("S0", "tau_species_ss", "tau_absorber_ss", "S"):
        print(e, k, "cpu-vs-stub", rel_err(a[k], b[k], 1e-30), "cpu-vs-GpuModel", rel_err(a[k], c[k], 1e-30), "stub-vs-GpuModel", rel_err(b[k], c[k], 1e-30))
    ta = cpu.arrays(e); tg = binding.define_singlet_tables(scn, e, binding.F32)
    for k in ("T_ratio", "density", "dtau_species", "dtau_absorber", "T_ratio_pt"):
        print("   table", k, np.array_equal(ta[k], tg[k]), rel_err(ta[k], tg[k]))
