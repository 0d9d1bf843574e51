"""which phase of a D-size set overlaps across contexts (threads) on one GPU: calls/s of each phase alone with 1, 2, 4 threads"""
import importlib, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
scn = synth.make_scenario(40, 20, 7, 12, n_em=2)
locs, dirs = synth.random_los(10000)
NT = 4
ctxs = []
for t in range(NT):
    ctx = binding.Context(0, binding.F64)
    g = ctx.make_grid(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, scn.rb, scn.szamethod, scn.raymethod)
    ctx.set_grid(g)
    tabs = [binding.define_singlet_tables(scn, e) for e in range(2)]
    for e in range(2):
        b, Tr, s, gf = (float(x) for x in scn.em_scalars[e]); ctx.set_singlet(e, 2, b, Tr, s, gf, tabs[e])
    ctx.generate_S()
    los = ctx.los_from_MSO(locs, dirs)
    ctx.los_upload(los)
    ctx.brightness_resident(10)
    ctxs.append((ctx, g, tabs, los))

def phase(name):
    def f(c):
        ctx, g, tabs, los = c
        if name == "set_grid": ctx.set_grid(g); [ctx.set_singlet(e, 2, *(float(x) for x in scn.em_scalars[e]), tabs[e]) for e in range(2)]
        elif name == "influence": ctx.influence(0, scn.n_vox)
        elif name == "solve": ctx.solve()
        elif name == "brightness_resident": ctx.brightness_resident(10)
        elif name == "brightness_host": ctx.brightness(los, 10)
    return f

for name in ("set_grid", "influence", "solve", "brightness_resident", "brightness_host"):
    if name != "set_grid":
        for c in ctxs:      # state needed by later phases
            pass
    res = []
    for nt in (1, 2, 4):
        reps = 60
        def work(c):
            f = phase(name)
            for _ in range(reps): f(c)
        th = [threading.Thread(target=work, args=(ctxs[i],)) for i in range(nt)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        res.append(f"{nt} thr: {dt / reps * 1e3:.3f} ms/round ({nt * reps / dt:.0f} calls/s)")
    print(f"{name:20s}", " | ".join(res), flush=True)
    if name == "set_grid":
        for c in ctxs: c[0].generate_S(); c[0].los_upload(c[3])

print("--- solve: device time per call (events on the ctx stream) vs wall, 1 / 2 / 4 threads")
for nt in (1, 2, 4):
    reps = 40
    out = [None] * nt
    def work(i):
        ctx = ctxs[i][0]
        ms = []
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.solve(); ms.append(ctx.kernel_ms(binding.PH_SOLVE)[0])
        out[i] = (np.median(ms), (time.perf_counter() - t0) / reps * 1e3)
    th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
    for t in th: t.start()
    for t in th: t.join()
    print(nt, "threads: device ms %s  wall ms/call %s" % ([round(o[0], 3) for o in out], [round(o[1], 3) for o in out]), flush=True)
os.environ["B200RT_SOLVE_TRACE"] = "1"
ctxs[0][0].solve()
