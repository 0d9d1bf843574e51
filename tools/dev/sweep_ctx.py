import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import extra_bench
for nlos in (10000, 500):
    for c in (1, 2, 4, 8):
        r = extra_bench.sweep(256, nlos, c, 1)
        print("n_los", nlos, "contexts", c, "sets/s", round(r["sweep_sets_per_s"], 1), "ms/set", round(1e3 / r["sweep_sets_per_s"], 3), flush=True)
