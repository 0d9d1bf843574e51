"""Generate golden input/output vectors from the REFERENCE's own hot-path source
(oracle/_ref, built in place from /root/reference by oracle/Makefile).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The .npz files it writes are committed; on the GPU box (no /root/reference) the oracle
and the CUDA path are checked against them.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
from oracle import refbind  # noqa: E402

CASES = [("g8x6x4x4", (8, 6, 4, 4), "f64"), ("g8x6x4x4", (8, 6, 4, 4), "f32"), ("g12x8x5x6", (12, 8, 5, 6), "f64")]


def main():
    for name, shape, prec in CASES:
        scn = synth.make_scenario(*shape, n_em=2, sza_T_contrast=0.1)
        R = refbind.RefModel(scn, prec)
        out = dict(shape=np.array(shape), rb=scn.rb, rexo=scn.rexo, szamethod=scn.szamethod, raymethod=scn.raymethod,
                   em_scalars=scn.em_scalars, abs_sigma=scn.abs_sigma, vox_in=scn.vox_in)
        for k, v in R.grid().items():
            out["grid_" + k] = v
        ln, eb, ent, dist = R.traverse_voxel_rays()
        out.update(vr_len=ln, vr_eb=eb, vr_ent=ent, vr_dist=dist)
        _, nsteps = R.build_rows()
        out["n_steps"] = nsteps
        for e in range(2):
            out[f"K{e}"] = R.K(e)
            for k, v in R.arrays(e).items():
                out[f"arr{e}_{k}"] = v
        R.solve()
        for e in range(2):
            for k, v in R.vectors(e).items():
                out[f"vec{e}_{k}"] = v
        locs_a, dirs_a = synth.fake_image(30 * synth.rMars, 30, 16)
        locs_b, dirs_b = synth.random_los(300, seed=7)
        locs = np.concatenate([locs_a, locs_b])
        dirs = np.concatenate([dirs_a, dirs_b])
        ln, eb, ent, dist, rs = R.traverse_los(locs, dirs)
        out.update(los_loc=locs, los_dir=dirs, los_len=ln, los_eb=eb, los_ent=ent, los_dist=dist, los_rayscal=rs)
        for nsub in (10, 0):
            _, b = R.brightness(locs, dirs, nsub)
            out[f"brightness_nsub{nsub}"] = b
        path = os.path.join(HERE, f"{name}_{prec}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
