"""Ad-hoc GPU-vs-oracle comparison (development aid; the real checks are tests/ -m gpu)."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d_planetary_rt_model_b200")
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
from oracle import oraclebind  # noqa: E402


def rel(a, b):
    m = (a != 0) | (b != 0)
    if not m.any():
        return 0.0
    return float((np.abs(a - b)[m] / np.maximum(np.abs(a[m]), np.abs(b[m]))).max())


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
    shape = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (40, 20, 7, 12)
    scn = synth.make_scenario(*shape, n_em=2, sza_T_contrast=0.1)
    O = oraclebind.OracleModel(scn, prec)
    G = binding.GpuModel(scn, prec)
    go, gg = O.grid(), G.grid()
    for k in go:
        print("grid", k, np.array_equal(go[k], gg[k]))
    a = O.traverse_voxel_rays()
    t0 = time.time()
    b = G.traverse_voxel_rays()
    print("traverse gpu s", time.time() - t0)
    print("trav len", np.array_equal(a[0], b[0]), "eb", np.array_equal(a[1], b[1]), "ent", np.array_equal(a[2], b[2]),
          "dist bits", np.array_equal(a[3].view(np.int64), b[3].view(np.int64)), "entries", len(a[2]), len(b[2]))
    if not np.array_equal(a[0], b[0]):
        bad = np.nonzero(a[0] != b[0])[0]
        print("  first bad rays", bad[:10], a[0][bad[:10]], b[0][bad[:10]])
    to, nso = O.build_rows()
    tg, nsg = G.build_rows()
    print("steps", nso, nsg, "oracle s", to, "gpu kernel s", tg)
    for e in range(scn.n_em):
        Ko, Kg = O.K(e), G.K(e)
        print("K", e, "max rel", rel(Ko, Kg), "pattern", np.array_equal(Ko != 0, Kg != 0))
        vo, vg = O.vectors(e), G.vectors(e, want_S=False)
        for k in ("S0", "tau_species_ss", "tau_absorber_ss"):
            print("  ", k, rel(vo[k], vg[k]))
        d = np.abs(vo["S0"] - vg["S0"]) / np.maximum(np.abs(vo["S0"]), 1e-300)
        w = np.argsort(d)[-5:]
        print("   worst S0 voxels", w, vo["S0"][w], vg["S0"][w], vo["tau_species_ss"][w])
    ro = O.solve()
    rg = G.solve()
    print("residuals oracle", ro, "gpu", rg, "solve ms", G.ctx.kernel_ms(binding.PH_SOLVE))
    for e in range(scn.n_em):
        print("S", e, rel(O.vectors(e)["S"], G.vectors(e)["S"]))
    for name, (locs, dirs) in dict(outside=synth.fake_image(30 * synth.rMars, 30, 60), inside=synth.random_los(4000)).items():
        a = O.traverse_los(locs, dirs)
        b = G.traverse_los(locs, dirs)
        print(name, "los trav len", np.array_equal(a[0], b[0]), "eb", np.array_equal(a[1], b[1]), "ent", np.array_equal(a[2], b[2]),
              "dist bits", np.array_equal(a[3].view(np.int64), b[3].view(np.int64)))
        for ns in (10, 0):
            to, bo = O.brightness(locs, dirs, ns)
            tg, bg = G.brightness(locs, dirs, ns)
            print("  brightness nsub", ns, "max rel", [rel(bo[:, q], bg[:, q]) for q in range(4)], "t", to, tg)


if __name__ == "__main__":
    main()
