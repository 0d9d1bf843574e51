"""per-stage wall time of one D-size parameter set (40x20 grid, 7x12 rays, 2 emissions, 1e4 LOS) on one context"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
scn = synth.make_scenario(40, 20, 7, 12, n_em=2)
ctx = binding.Context(0, binding.F64)
g = ctx.make_grid(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, scn.rb, scn.szamethod, scn.raymethod)
tabs = [binding.define_singlet_tables(scn, e) for e in range(2)]
locs, dirs = synth.random_los(10000)
los = ctx.los_from_MSO(locs, dirs)
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for it in range(30):
    t0 = time.perf_counter(); ctx.set_grid(g); tick("set_grid", t0)
    t0 = time.perf_counter()
    for e in range(2):
        b, Tr, s, gf = (float(x) for x in scn.em_scalars[e]); ctx.set_singlet(e, 2, b, Tr, s, gf, tabs[e])
    tick("set_singlet x2", t0)
    t0 = time.perf_counter(); ctx.influence(0, scn.n_vox); tick("influence", t0)
    k = {p: ctx.kernel_ms(p)[0] for p in (binding.PH_TRAVERSE, binding.PH_INFLUENCE)}
    t0 = time.perf_counter(); ctx.solve(); tick("solve", t0)
    ks = ctx.kernel_ms(binding.PH_SOLVE)[0]
    t0 = time.perf_counter(); out = ctx.brightness(los, 10); tick("brightness(host)", t0)
    kb = {p: ctx.kernel_ms(p)[0] for p in (binding.PH_TRAVERSE, binding.PH_BRIGHTNESS)}
for k_, v in T.items():
    print(f"{k_:18s} {np.median(v[5:]):7.3f} ms")
print("kernel ms: traverse %.3f march %.3f solve %.3f los-traverse %.3f brightness %.3f" % (k[0], k[1], ks, kb[0], kb[3]))
print("sum of stages %.3f ms" % sum(np.median(v[5:]) for v in T.values()))
