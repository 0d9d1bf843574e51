"""Golden vectors for the plane-parallel grid <40, 7> (observation_fit.hpp:48-50), produced by the
REFERENCE's own source (oracle/_ref, plane_parallel_grid + singlet_CFR + RT_grid, built in place).
Run where /root/reference exists:  python tests/golden/make_pp_golden.py"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
from oracle import refbind  # noqa: E402


def main():
    for prec in ("f64", "f32"):
        scn = synth.make_scenario_pp(40, 7, n_em=2)
        R = refbind.RefModel(scn, prec)
        out = dict(rb=scn.rb, rexo=scn.rexo, vox_in=scn.vox_in, em_scalars=scn.em_scalars, abs_sigma=scn.abs_sigma)
        g = R.grid()
        for k in ("pts_radii", "ray_theta", "ray_domega"):
            out["grid_" + k] = g[k]
        ln, eb, ent, dist = R.traverse_voxel_rays()
        out.update(vr_len=ln, vr_eb=eb, vr_ent=ent, vr_dist=dist)
        _, out["n_steps"] = R.build_rows()
        for e in range(2):
            out[f"K{e}"] = R.K(e)
        R.solve()
        for e in range(2):
            for k, v in R.vectors(e).items():
                out[f"vec{e}_{k}"] = v
        path = os.path.join(HERE, f"pp40x7_{prec}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
