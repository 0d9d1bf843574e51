"""Golden vectors for the multiplet CFR emissions (O I 102.6 nm, H Lyman multiplet, H Lyman singlet-as-multiplet),
produced by the REFERENCE's own source (oracle/_ref/libref_mult_*.so: multiplet_CFR_emission.hpp, O_1026.hpp,
H_lyman_multiplet.hpp, H_lyman_multiplet_test.hpp built in place).
Run where /root/reference exists:  python tests/golden/make_mult_golden.py"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
from oracle import multbind  # noqa: E402

CASES = [(0, "f64"), (0, "f32"), (1, "f64"), (1, "f32"), (2, "f64")]
SHAPE = (8, 6, 4, 4)


def main():
    for kind, prec in CASES:
        scn = synth.make_multiplet_scenario(kind, *SHAPE, sza_T_contrast=0.1)
        R = multbind.RefMultiplet(scn, prec)
        out = dict(kind=kind, shape=np.array(SHAPE), rb=scn.rb, rexo=scn.rexo, szamethod=scn.szamethod,
                   raymethod=scn.raymethod, solar=scn.solar, vox_in=scn.vox_in)
        for k, v in R.constants().items():
            out["const_" + k] = v
        for k, v in R.arrays().items():
            out["arr_" + k] = v
        _, out["n_steps"] = R.build_rows()
        out["K"] = R.K()
        R.solve()
        for k, v in R.vectors().items():
            out["vec_" + k] = v
        locs_a, dirs_a = synth.fake_image(30 * synth.rMars, 30, 12)
        locs_b, dirs_b = synth.random_los(200, seed=11)
        locs, dirs = np.concatenate([locs_a, locs_b]), np.concatenate([dirs_a, dirs_b])
        out.update(los_loc=locs, los_dir=dirs)
        for nsub in (10, 0):
            for k, v in R.brightness(locs, dirs, nsub).items():
                out[f"b{nsub}_{k}"] = v
        path = os.path.join(HERE, f"mult{kind}_{prec}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
