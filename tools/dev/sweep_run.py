"""Development aid: sweep throughput against contexts per GPU (python tools/dev/sweep_run.py)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import extra_bench
for c in (1, 2, 4):
    r = extra_bench.sweep(512, 10000, c, 1)
    print(c, "contexts:", round(r["sweep_seconds"], 3), "s", round(r["sweep_sets_per_s"], 1), "sets/s", flush=True)
