"""Generates tests/golden/iph_real_table.npz: the reference's Quemerais table
(src/quemerais_IPH_model/fsm99td12v20t80) parsed by the oracle, plus the oracle's output for a
seeded set of sky directions -- so that the GPU box (no /root/reference) checks the device kernel on
the real model.  Run in the build container:  python tests/golden/make_iph_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
from oracle import iphbind  # noqa: E402

O = iphbind.IphOracle(fname=iphbind.REF_TABLE)
tab = O.table()
ra, dec = synth.random_sky(600, seed=11)
g = synth.lyman_alpha_typical_g_factor
kR = O.model(g, synth.MARS_ECLIPTIC_POS, ra, dec)
out = os.path.join(ROOT, "tests", "golden", "iph_real_table.npz")
np.savez_compressed(out, ra=ra, dec=dec, g_lya=g, marspos=np.array(synth.MARS_ECLIPTIC_POS), kR=kR,
                    **{"tab_" + k: np.asarray(v) for k, v in tab.items()})
print("wrote", out, os.path.getsize(out), "bytes; kR range", kR.min(), kR.max())
