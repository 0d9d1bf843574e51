import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = [("g8x6x4x4", "f64"), ("g8x6x4x4", "f32"), ("g12x8x5x6", "f64")]


def load_golden(synth, name, prec):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}_{prec}.npz"))
    n_rb, n_sb, n_th, n_ph = (int(x) for x in z["shape"])
    scn = synth.Scenario(n_rb, n_sb, n_th, n_ph, z["rb"], float(z["rexo"]), int(z["szamethod"]), int(z["raymethod"]),
                         z["em_scalars"], z["abs_sigma"], z["vox_in"])
    return scn, z
