"""Development aid: where do the float-Real outputs of the CUDA path and of the oracle differ most?"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
from oracle import oraclebind  # noqa: E402


def worst(a, b, k=4):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    d = np.abs(a - b) / den
    d[(a == 0) & (b == 0)] = 0
    w = np.argsort(d)[-k:][::-1]
    return [(int(i), float(d[i]), float(a[i]), float(b[i])) for i in w]


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
    for shape, los_sets in (((12, 8, 5, 6), [synth.fake_image(30 * synth.rMars, 30, 24), synth.random_los(800)]),
                            ((40, 20, 7, 12), [synth.fake_image(30 * synth.rMars, 30, 40), synth.random_los(3000)])):
        scn = synth.make_scenario(*shape, n_em=2, sza_T_contrast=0.1 if shape[0] == 12 else 0.0)
        O = oraclebind.OracleModel(scn, prec)
        G = binding.GpuModel(scn, prec)
        O.build_rows(); G.build_rows()
        for e in range(2):
            print(shape, "K", e, worst(O.K(e).ravel(), G.K(e).ravel(), 2))
        O.solve(); G.solve()
        for e in range(2):
            So = O.vectors(e)["S"]
            print(shape, "S", e, worst(So, G.vectors(e)["S"], 2))
            G.set_sourcefn(e, So)
        for li, (locs, dirs) in enumerate(los_sets):
            for nsub in (10, 0, 4):
                _, bo = O.brightness(locs, dirs, nsub)
                _, bg = G.brightness(locs, dirs, nsub)
                for q in range(4):
                    for e in range(2):
                        w = worst(bo[e, q], bg[e, q], 3)
                        if w[0][1] > 3e-5:
                            print(shape, "los", li, "nsub", nsub, "q", q, "e", e, w, "max", float(np.abs(bo[e, q]).max()))


if __name__ == "__main__":
    main()
