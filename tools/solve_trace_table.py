"""Tabulate the last solve of a B200RT_SOLVE_TRACE log: python tools/solve_trace_table.py gpurun_out/solve_trace.log"""
import sys
lines = [l.split() for l in open(sys.argv[1]) if l.startswith('solve-trace')]
runs, cur = [], []
for l in lines:
    if l[1] == 'start' and cur:
        runs.append(cur); cur = []
    k = l[2].replace('K=', '') or l[3]
    cur.append((l[1], int(k), float(l[-2])))
runs.append(cur)
d = {(w, k): t for w, k, t in runs[-1]}
print("factor_end", d[('factor_end', 0)], "backsolve_end", d[('backsolve_end', 0)])
print("  K  chain_b chain_e    ui_b    ui_e   uii_e | chain_len  ui_len uii_len")
f = lambda x: f"{x:7.3f}" if x is not None else "   -   "
sub = lambda a, b: (a - b) if (a is not None and b is not None) else None
nK = max(k for (w, k) in d if w == 'chain_begin') + 1
for K in range(0, nK, max(1, nK // 16)):
    cb, ce, ub, ue, ve = (d.get((w, K)) for w in ('chain_begin', 'chain_end', 'ui_begin', 'ui_end', 'uii_end'))
    print(f"{K:3d} {f(cb)} {f(ce)} {f(ub)} {f(ue)} {f(ve)} |  {f(sub(ce, cb))} {f(sub(ue, ub))} {f(sub(ve, ue))}")
