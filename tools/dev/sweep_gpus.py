"""sweep rate against (GPUs, contexts per GPU) in one process: python tools/dev/sweep_gpus.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import extra_bench
import importlib
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
n = binding.load().b200rt_device_count()
print("devices", n, "cpus", os.cpu_count(), flush=True)
for g in sorted({n, max(1, n // 2)}, reverse=True):
    for c in (1, 2, 3, 4):
        r = extra_bench.sweep(1024, 10000, c, g)
        print("gpus", g, "contexts", c, "sets/s best", round(r["sweep_sets_per_s"], 1), [round(x) for x in r["sweep_sets_per_s_runs"]], flush=True)
