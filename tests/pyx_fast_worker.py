"""Run in a subprocess by tests/test_pyx_fast.py: the buffer-protocol / nogil binding (host/py_corona_sim_b200.pyx, built
by oracle/build_pyx.py into oracle/_ref/py_corona_sim_fast/) against the reference's own class in the same module."""
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "py_corona_sim_fast"))
mod = importlib.import_module("py_corona_sim_gpu")
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")

out = {}
n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
locs, dirs = synth.random_los(n_big)
ra, dec = synth.random_sky(n_big)

slow, fast = mod.Pyobservation_fit(), mod.Pyobservation_fit_b200()
n_small = 20000
t0 = time.perf_counter(); slow.add_observation(locs[:n_small], dirs[:n_small]); t_slow = time.perf_counter() - t0
t0 = time.perf_counter(); fast.add_observation(locs, dirs); t_fast = time.perf_counter() - t0
out["add_observation_s_per_1e6_reference_binding"] = t_slow / n_small * 1e6
out["add_observation_s_per_1e6_fast_binding"] = t_fast / n_big * 1e6

# identical results on the same lines of sight
fast.add_observation(locs[:n_small], dirs[:n_small])
for P in (slow, fast):
    P.generate_source_function(5e5, 200.0)
bs, bf = np.asarray(slow.brightness()), fast.brightness()
out["brightness_shape"] = list(bf.shape)
out["brightness_max_rel"] = float(np.max(np.abs(bs - bf) / np.maximum(np.abs(bs), 1e-300)))
out["col_dens_max_rel"] = float(np.max(np.abs(np.asarray(slow.species_col_dens()) - fast.species_col_dens())
                                       / np.maximum(np.abs(fast.species_col_dens()), 1e-300)))
slow.add_observation_ra_dec(np.array(synth.MARS_ECLIPTIC_POS), ra[:n_small], dec[:n_small])
fast.add_observation_ra_dec(synth.MARS_ECLIPTIC_POS, ra[:n_small], dec[:n_small])      # a tuple: any buffer / sequence
out["iph_equal"] = bool(np.array_equal(np.asarray(slow.iph_brightness_unextincted()), fast.iph_brightness_unextincted()))
out["iph_shape"] = list(fast.iph_brightness_unextincted().shape)

# the GIL is released: a Python thread makes progress while generate_source_function + brightness run
fast.simulate_iph(False)              # the IPH coordinates above belong to the smaller set
fast.add_observation(locs, dirs)
ticks, stop = [0], [False]


def spin():
    while not stop[0]:
        ticks[0] += 1
        time.sleep(0.0005)


th = threading.Thread(target=spin)
th.start()
t0 = time.perf_counter()
fast.generate_source_function(5e5, 250.0)
b = fast.brightness()
dt = time.perf_counter() - t0
stop[0] = True
th.join()
out["gil_free_call_s"] = dt
out["ticks_during_call"] = ticks[0]
out["ticks_expected_if_released"] = dt / 0.0005 * 0.3
out["finite"] = bool(np.isfinite(b).all())
print("RESULT " + json.dumps(out))
