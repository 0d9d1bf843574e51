import sys, os
os.environ["B200RT_BATCH_PROFILE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import extra_bench, importlib
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
n = binding.load().b200rt_device_count()
for g, c in ((1, 1), (n, 1), (n, 2)) if n > 1 else ((1, 1), (1, 4)):
    print("gpus", g, "contexts", c, flush=True)
    extra_bench.sweep(1024, 10000, c, g)
