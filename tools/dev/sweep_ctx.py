"""sweep rate on one GPU against the number of contexts (worker threads) sharing it: 990 sets x 1e4 lines of sight"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import extra_bench
for c in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (1, 2, 4, 6, 8):
    r = extra_bench.sweep(1024, 10000, c, 1)
    print("contexts", c, "sets/s best", round(r["sweep_sets_per_s"], 1), "runs", [round(x) for x in r["sweep_sets_per_s_runs"]], flush=True)
