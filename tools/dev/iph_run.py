"""Development aid: one IPH run (python tools/dev/iph_run.py [n_los] [sorted])"""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import extra_bench
print(json.dumps(extra_bench.iph(int(sys.argv[1]) if len(sys.argv) > 1 else 1000000)))
